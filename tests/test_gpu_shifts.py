"""The whole path on the GPU: prepare_search_groups -> calculate_shifts (batched probes) -> grouping
heuristics, against the reference's golden run and against the same host logic driven by the CPU
oracle on identical inputs."""
import numpy as np
import pytest

from sushi_b200 import WavStream, synth, grouping
from sushi_b200.events import ScriptEvent
from sushi_b200.grouping import prepare_search_groups
from sushi_b200.shifts import calculate_shifts
from tests.helpers import oracle_stream_from_pcm
from tests.test_shifts_host import scenario_inputs

pytestmark = pytest.mark.gpu
SAMPLE = 1.0 / 12000 + 1e-9


@pytest.mark.parametrize('name', ['const', 'jump', 'rewind'])
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_gpu_solver_matches_reference_golden(gpu_lib, golden_shifts, name, stype):
    g = golden_shifts
    src_pcm, dst_pcm, ev, params = scenario_inputs(g, name)
    src = WavStream.from_pcm(src_pcm, 12000, sample_type=stype)
    dst = WavStream.from_pcm(dst_pcm, 12000, sample_type=stype)
    events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(ev)]
    groups = prepare_search_groups(events, src.duration_seconds, [], 0.417, 0.417)
    calculate_shifts(src, dst, groups, float(params[0]), float(params[1]), int(params[2]))
    want = g['{0}_{1}_result'.format(name, stype)]
    got = np.array([[e.shift, e.diff, (e._link.source_index if e.linked else -1)] for e in events])
    assert np.array_equal(got[:, 2], want[:, 2])                       # same link structure
    assert np.abs(got[:, 0] - want[:, 0]).max() <= SAMPLE               # shifts within one sample
    assert np.abs(got[:, 1] - want[:, 1]).max() <= 1e-5                 # diffs within 1e-5


def test_grouped_mode_config4_miniature(gpu_lib):
    """BASELINE config 4 in miniature: chapters with their own shift, events grouped by chapter,
    the post-processing heuristics applied; GPU run == CPU-oracle run event by event."""
    rng = np.random.default_rng(44)
    chapters = [0.0, 40.0, 95.0, 150.0, 210.0]
    shifts = [0.5, -1.25, 3.0, 3.0, -0.4]
    dur = 260.0
    src_pcm, dst_pcm = synth.make_pair(dur, 44, list(zip(chapters, shifts)))
    starts, ends = synth.make_events(60, dur - 10.0, 44, 0.8, 3.0, 2.0)

    def run(make_stream):
        src, dst = make_stream(src_pcm), make_stream(dst_pcm)
        events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(zip(starts, ends))]
        groups = prepare_search_groups(events, src.duration_seconds, chapters, 0.417, 0.417)
        calculate_shifts(src, dst, groups, 10.0, 30.0, 5)
        ev = [e for e in events if not e.linked]
        by_chapter = grouping.groups_from_chapters(ev, chapters)
        for grp in by_chapter:
            grouping.fix_near_borders(grp)
            grouping.smooth_events([e for e in grp if not e.linked], 3)
        by_chapter = grouping.split_broken_groups(by_chapter)
        for grp in by_chapter:
            grouping.average_shifts(grp)
        return np.array([[e.shift, e.diff] for e in events])

    gpu = run(lambda pcm: WavStream.from_pcm(pcm, 12000))
    cpu = run(lambda pcm: oracle_stream_from_pcm(pcm, 12000, 1, 12000, 'uint8'))
    assert np.abs(gpu[:, 0] - cpu[:, 0]).max() <= SAMPLE
    assert np.abs(gpu[:, 1] - cpu[:, 1]).max() <= 1e-5
    # and the known answer: every event carries its chapter's shift
    mid = (starts + ends) / 2
    truth = np.array(shifts)[np.searchsorted(chapters, mid, side='right') - 1]
    inside = np.array([not any(a < c < b + 0.01 for c in chapters[1:]) for a, b in zip(starts, ends)])
    assert np.abs(gpu[inside, 0] - truth[inside]).max() <= 0.011


def test_find_substream_many_equals_singles(gpu_lib, golden_matcher):
    from tests.helpers import oracle_stream_from_pcm as mk
    rs = mk(golden_matcher['src_pcm'], 12000, 1, 12000, 'uint8')
    rd = mk(golden_matcher['dst_pcm'], 12000, 1, 12000, 'uint8')
    src = WavStream.from_array(rs.data, 12000, rs.padding_size, rs.sample_count)
    dst = WavStream.from_array(rd.data, 12000, rd.padding_size, rd.sample_count)
    tv = src.get_substream(6.1, 9.05)
    half = len(tv[0]) // 2
    qs = [(tv, 7.6, 10.0), (tv[:, :half], 7.6, 10.0), (tv[:, half:], 7.6 + half / 12000.0, 10.0)]
    many = dst.find_substream_many(qs)
    for q, (d, t) in zip(qs, many):
        d1, t1 = dst.find_substream(*q)
        assert d == d1 and t == t1
    # a detached copy cannot be located: falls back to per-call uploads, same answers
    many2 = dst.find_substream_many([(tv.copy(), 7.6, 10.0)] + qs[1:])
    assert many2[0] == many[0]


@pytest.mark.parametrize('name', ['const', 'jump', 'rewind'])
def test_speculative_solver_is_bit_identical_to_live_calls(gpu_lib, golden_shifts, name):
    """Answering the fast-path searches from precomputed curves changes nothing: same shifts, same
    diffs (bit for bit), same links, as the run that issues every find_substream live."""
    g = golden_shifts
    src_pcm, dst_pcm, ev, params = scenario_inputs(g, name)
    out = []
    for speculative in (False, True):
        src = WavStream.from_pcm(src_pcm, 12000)
        dst = WavStream.from_pcm(dst_pcm, 12000)
        events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(ev)]
        groups = prepare_search_groups(events, src.duration_seconds, [], 0.417, 0.417)
        calculate_shifts(src, dst, groups, float(params[0]), float(params[1]), int(params[2]), speculative=speculative)
        out.append([(e.shift, float(e.diff), e._link.source_index if e.linked else -1) for e in events])
        if speculative:
            assert dst.__dict__.get('_curve_cache')           # the speculation path was exercised
    assert out[0] == out[1]


def test_cached_curve_answers_equal_live_answers(gpu_lib, golden_matcher):
    rs = oracle_stream_from_pcm(golden_matcher['src_pcm'], 12000, 1, 12000, 'uint8')
    rd = oracle_stream_from_pcm(golden_matcher['dst_pcm'], 12000, 1, 12000, 'uint8')
    src = WavStream.from_array(rs.data, 12000, rs.padding_size, rs.sample_count)
    dst = WavStream.from_array(rd.data, 12000, rd.padding_size, rd.sample_count)
    groups = [[ScriptEvent(i, a, a + 1.3)] for i, a in enumerate(np.arange(1.0, 19.0, 1.7))]
    live = [dst.find_substream(src.get_substream(grp[0].start, grp[0].end), grp[0].start + 1.5 + 0.003 * i, 1.5)
            for i, grp in enumerate(groups)]
    dst.speculate_fast_path(src, groups, 0, 1.5, 1.5)
    cached = [dst.find_substream(src.get_substream(grp[0].start, grp[0].end), grp[0].start + 1.5 + 0.003 * i, 1.5)
              for i, grp in enumerate(groups)]
    assert live == cached
    # a range outside the cached span falls back to a live call and still agrees with the oracle
    d, t = dst.find_substream(src.get_substream(1.0, 2.3), 6.0, 1.5)
    d_ref, t_ref = rd.find_substream(rs.get_substream(1.0, 2.3), 6.0, 1.5)
    assert abs(float(d) - float(d_ref)) <= 1e-5 and abs(t - t_ref) <= SAMPLE


def test_shift_script_end_to_end(gpu_lib, tmp_path):
    """WAV + ASS in, shifted ASS out (sushi.py:660-726 without demux/keyframes): stereo 48 kHz files go
    through the GPU loader, the events through the solver and the heuristics, and every dialogue line
    lands on the known shift; comments follow the line they are linked to."""
    import wave
    from sushi_b200 import shift_script, AssScript
    dur, shift = 70.0, 2.25
    src12, dst12 = synth.make_pair(dur, 31, shift)
    for name, pcm in (('src.wav', src12), ('dst.wav', dst12)):
        up = np.repeat(pcm, 4)                                   # 48 kHz whose nearest-resample is the 12 kHz signal
        st = np.stack([up, up], 1)
        with wave.open(str(tmp_path / name), 'wb') as w:
            w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(st.tobytes())
    starts, ends = synth.make_events(24, dur - 8.0, 31, 0.8, 3.0, 1.5)
    from sushi_b200.common import format_time
    lines = ['[Script Info]', 'Title: t', '', '[V4+ Styles]', AssScript.STYLES_FORMAT,
             'Style: Default,Arial,20,&H00FFFFFF,&H000000FF,&H00000000,&H00000000,0,0,0,0,100,100,0,0,1,2,2,2,10,10,10,1',
             '', '[Events]', AssScript.EVENTS_FORMAT]
    for i, (a, b) in enumerate(zip(starts, ends)):
        kind = 'Comment' if i == 5 else 'Dialogue'
        lines.append('{0}: 0,{1},{2},Default,,0,0,0,,line {3}, with a comma'.format(kind, format_time(a), format_time(b), i))
    (tmp_path / 'in.ass').write_text('\n'.join(lines), encoding='utf-8')
    script, groups = shift_script(str(tmp_path / 'src.wav'), str(tmp_path / 'dst.wav'), str(tmp_path / 'in.ass'),
                                  str(tmp_path / 'out.ass'))
    out = AssScript.from_file(str(tmp_path / 'out.ass'))
    assert len(out.events) == len(starts) and len(groups) == 1
    for e, a, b in zip(out.events, starts, ends):
        # ASS stores centiseconds: the written time is the shifted time rounded to 0.01 s
        assert abs(e.start - (round(a * 100) / 100 + shift)) <= 0.011 and abs(e.end - (round(b * 100) / 100 + shift)) <= 0.011
    assert out.events[5].is_comment and out.events[0].text == 'line 0, with a comma'
