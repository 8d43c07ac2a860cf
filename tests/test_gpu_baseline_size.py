"""GPU against the CPU oracle AT BASELINE SIZE (BASELINE.json configs[2], [3], [4]): 2 x 90-minute streams
(65 M-sample running sums, 3 970 block-spectrum rows), +-120 s / +-300 s / +-600 s search windows.

Config 2 has its own checks in test_gpu_matcher.py; here the big geometries get the same bar on `diff` and on
the shift -- north_star's tolerances, written out: shift within +-1 destination sample (1/12000 s), diff within
1e-5 of the reference's cv2.matchTemplate path -- on a sample of events the oracle finishes in seconds, plus
size-independent properties over the whole batch (the known shift comes back for all 10 000 events; the
sharded path returns what the single call returns)."""
import numpy as np
import pytest

from sushi_b200 import WavStream, synth, grouping
from sushi_b200.events import ScriptEvent
from sushi_b200.grouping import prepare_search_groups
from sushi_b200.shifts import calculate_shifts
from tests.helpers import oracle_stream_from_pcm

pytestmark = pytest.mark.gpu
SAMPLE = 1.0 / 12000 + 1e-9
DUR = 5400.0


@pytest.fixture(scope='module')
def ninety(gpu_lib):
    """The config-3 inputs of bench.py: +1.5 s constant shift, seed 2; oracle streams and GPU streams over the
    same normalised arrays."""
    src_pcm, dst_pcm = synth.make_pair(DUR, 2, 1.5)
    rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, 'uint8')
    rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, 'uint8')
    src = WavStream.from_array(rs.data, 12000, rs.padding_size, rs.sample_count)
    dst = WavStream.from_array(rd.data, 12000, rd.padding_size, rd.sample_count)
    yield rs, rd, src, dst
    src.close()
    dst.close()


def test_config3_sample_of_events_against_the_oracle(ninety):
    rs, rd, src, dst = ninety
    starts, ends = synth.make_events(10000, DUR, 2, 1.0, 4.0)
    win = np.full(len(starts), 120.0)
    diffs, times = dst.find_substream_batch(src, starts, ends, starts, win)
    # the known answer, all 10 000 events
    ok = ends + 1.5 < DUR
    assert ok.sum() > 9990
    assert np.abs((times - starts)[ok] - 1.5).max() <= SAMPLE
    # the oracle (reference find_substream over cv2) on 24 events spread over the stream, first and last included
    # (their windows are clipped at the stream's ends, wav.py:178-179)
    pick = np.unique(np.concatenate([[0, 1, len(starts) - 2, len(starts) - 1], np.linspace(0, len(starts) - 1, 20).astype(int)]))
    worst_d = worst_t = 0.0
    for q in pick:
        d_ref, t_ref = rd.find_substream(rs.get_substream(starts[q], ends[q]), starts[q], 120.0)
        worst_d = max(worst_d, abs(float(diffs[q]) - float(d_ref)))
        worst_t = max(worst_t, abs(times[q] - t_ref))
    assert worst_d <= 1e-5, worst_d
    assert worst_t <= SAMPLE, worst_t
    # single calls through the reference-shaped API give the batch's answers bit for bit
    for q in pick[:4]:
        d1, t1 = dst.find_substream(src.get_substream(starts[q], ends[q]), starts[q], 120.0)
        assert d1 == diffs[q] and t1 == times[q]


def test_config3_whole_curve_of_one_event(ninety):
    """All 2 880 001 lags of one +-120 s search in the middle of the 90-minute stream: the curve, not only its
    minimum, within 1e-5 of cv2's."""
    rs, rd, src, dst = ninety
    a, b = 2700.37, 2703.11
    toff, tlen, lag0, nlags, _ = dst.plan_queries(src, [a], [b], [a], [120.0])
    toff, tlen, lag0, nlags = int(toff[0]), int(tlen[0]), int(lag0[0]), int(nlags[0])
    assert nlags == 2880001
    got = dst.match_curve(src, toff, tlen, lag0, nlags)
    want = rd.match_curve(rs.data[:, toff:toff + tlen], lag0, nlags)
    assert np.abs(got - want).max() <= 1e-5
    assert abs(int(got.argmin()) - int(want.argmin())) <= 1
    d, i = dst.find_planned(src, [toff], [tlen], [lag0], [nlags])
    assert i[0] == int(got.argmin()) and d[0] == got.min()


@pytest.mark.parametrize('ev_len,window', [(0.5, 5.0), (30.0, 600.0), (30.0, 5.0), (0.5, 600.0)])
def test_config5_corners_against_the_oracle(ninety, ev_len, window):
    """The four corners of the config-5 sweep (event length 0.5 .. 30 s x window +-5 .. +-600 s, i.e. 6 000-sample
    templates of one partition up to 360 000 samples = 22 partitions -- the blocked multiply route -- and 120 001
    up to 14 400 001 lags)."""
    rs, rd, src, dst = ninety
    starts = np.array([700.25, 2650.5, 4600.75])
    ends = starts + ev_len
    diffs, times = dst.find_substream_batch(src, starts, ends, starts, np.full(3, window))
    for q in range(3):
        d_ref, t_ref = rd.find_substream(rs.get_substream(starts[q], ends[q]), starts[q], window)
        assert abs(float(diffs[q]) - float(d_ref)) <= 1e-5, (q, diffs[q], d_ref)
        assert abs(times[q] - t_ref) <= SAMPLE, (q, times[q], t_ref)
        assert abs((times[q] - starts[q]) - 1.5) <= SAMPLE


def test_sharded_matcher_on_one_gpu_equals_the_direct_call(ninety, gpu_lib):
    """parallel.ShardedMatcher (device buffers, planned from stream geometry alone, results through the padded
    gather layout) with a single rank: the answers of find_substream_batch, bit for bit."""
    from sushi_b200 import parallel
    rs, rd, src, dst = ninety
    starts, ends = synth.make_events(400, DUR, 9, 1.0, 4.0)
    win = np.full(len(starts), 120.0)
    want = dst.find_substream_batch(src, starts, ends, starts, win)
    be = parallel.DeviceBackend(gpu_lib)
    m = parallel.ShardedMatcher(parallel.SingleComm(be), be)
    m.set_streams(rs, rd)
    got = m.find_batch(starts, ends, starts, win)
    be.release()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def config4_inputs(n_chapters, per_chapter, seed=4, big_jumps=3):
    """BASELINE configs[3]: chapter groups with a slowly drifting per-chapter shift (steps <= 0.2 s), a dozen
    jumps of several seconds (re-acquired at the normal +-10 s window) and a few jumps beyond it (15 .. 40 s:
    these need the rewind to max_window = +-300 s, sushi.py:473-478)."""
    dur = n_chapters * 10.8
    rng = np.random.default_rng(seed)
    chapters = [i * (dur / n_chapters) for i in range(n_chapters)]
    steps = rng.uniform(-0.2, 0.2, n_chapters)
    some = rng.choice(np.arange(5, n_chapters - 5), min(12, n_chapters // 8) + big_jumps, replace=False)
    steps[some[big_jumps:]] += rng.uniform(-8, 8, len(some) - big_jumps)
    steps[some[:big_jumps]] += rng.uniform(15, 40, big_jumps) * rng.choice([-1, 1], big_jumps)
    shifts = np.round(np.cumsum(steps) * 12000) / 12000
    shifts -= np.round(shifts.mean() * 12000) / 12000
    src_pcm, dst_pcm = synth.make_pair(dur, seed, list(zip(chapters, shifts)))
    starts, ends = synth.make_events(n_chapters * per_chapter, dur, seed, 0.45, 0.9, 1.0)
    return dur, chapters, shifts, src_pcm, dst_pcm, starts, ends


def run_grouped_mode(make_stream, chapters, src_pcm, dst_pcm, starts, ends, window=10.0, max_window=300.0):
    """prepare_search_groups -> calculate_shifts -> the chapter post-processing chain of sushi.py:682-711."""
    src, dst = make_stream(src_pcm), make_stream(dst_pcm)
    events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(zip(starts, ends))]
    groups = prepare_search_groups(events, src.duration_seconds, chapters, 0.417, 0.417)
    calculate_shifts(src, dst, groups, window, max_window, 5)
    raw = np.array([[e.shift, e.diff] for e in events], np.float64)
    ev = [e for e in events if not e.linked]
    by_chapter = grouping.groups_from_chapters(ev, chapters)
    for grp in by_chapter:
        grouping.fix_near_borders(grp)
        grouping.smooth_events([e for e in grp if not e.linked], 3)
    by_chapter = grouping.split_broken_groups(by_chapter)
    for grp in by_chapter:
        grouping.average_shifts(grp)
    final = np.array([[e.shift, e.diff] for e in events], np.float64)
    links = [e._link.source_index if e.linked else -1 for e in events]
    return raw, final, links, len(groups)


def test_config4_grouped_mode_at_baseline_size(gpu_lib):
    """500 chapter groups x 20 events, piecewise drift, +-300 s max window: the GPU run against the same host
    logic on the oracle's streams (every find_substream through cv2), event by event -- the shifts calculate_shifts
    commits (+-1 sample, diff 1e-5), the link structure, and the final per-event shifts after fix_near_borders /
    smooth_events / split_broken_groups / average_shifts."""
    dur, chapters, shifts, src_pcm, dst_pcm, starts, ends = config4_inputs(500, 20)
    gpu = run_grouped_mode(lambda pcm: WavStream.from_pcm(pcm, 12000), chapters, src_pcm, dst_pcm, starts, ends)
    cpu = run_grouped_mode(lambda pcm: oracle_stream_from_pcm(pcm, 12000, 1, 12000, 'uint8'), chapters, src_pcm, dst_pcm, starts, ends)
    assert gpu[3] == cpu[3] and gpu[2] == cpu[2]
    assert np.abs(gpu[0][:, 0] - cpu[0][:, 0]).max() <= SAMPLE
    assert np.abs(gpu[0][:, 1] - cpu[0][:, 1]).max() <= 1e-5
    assert np.abs(gpu[1][:, 0] - cpu[1][:, 0]).max() <= SAMPLE
    assert np.abs(gpu[1][:, 1] - cpu[1][:, 1]).max() <= 1e-5
    # the known answer: an event that lies inside one chapter carries that chapter's shift
    mid = (starts + ends) / 2
    truth = shifts[np.searchsorted(chapters, mid, side='right') - 1]
    inside = np.array([np.searchsorted(chapters, a, side='right') == np.searchsorted(chapters, b + 0.01, side='right')
                       for a, b in zip(starts, ends)])
    err = np.abs(gpu[1][:, 0] - truth)
    assert np.mean(err[inside] <= 0.011) >= 0.93, np.mean(err[inside] <= 0.011)
