"""The shift solver and search-group preparation on the CPU, with the oracle's streams as the
matcher backend, against the reference's golden run of prepare_search_groups + calculate_shifts
(tests/golden/shifts.npz, produced by oracle/gen_golden.py from /root/reference/sushi.py)."""
import zlib

import numpy as np
import pytest

from sushi_b200 import synth
from sushi_b200.events import ScriptEvent
from sushi_b200.grouping import prepare_search_groups
from sushi_b200.shifts import calculate_shifts
from tests.helpers import oracle_stream_from_pcm


def scenario_inputs(g, name):
    dur, seed, count = g[name + '_gen']
    shift = [tuple(r) for r in g[name + '_shift']]
    src_pcm, dst_pcm = synth.make_pair(float(dur), int(seed), shift if len(shift) > 1 else shift[0][1])
    assert zlib.crc32(src_pcm.tobytes()) == int(g[name + '_pcm_crc'][0]), 'synthetic generator drifted'
    assert zlib.crc32(dst_pcm.tobytes()) == int(g[name + '_pcm_crc'][1])
    return src_pcm, dst_pcm, g[name + '_events'], g[name + '_params']


class TracingStream(object):
    """Wraps a stream and records every find_substream call as (offset, n, center, window, diff, time)."""

    def __init__(self, inner, src):
        self._inner, self._src, self.calls = inner, src, []
        self.sample_rate = inner.sample_rate

    @property
    def duration_seconds(self):
        return self._inner.duration_seconds

    def find_substream(self, pattern, center, window):
        d, t = self._inner.find_substream(pattern, center, window)
        off = (pattern.__array_interface__['data'][0] - self._src.data.__array_interface__['data'][0]) // self._src.data.itemsize
        self.calls.append((off, len(pattern[0]), center, window, float(d), t))
        return d, t


@pytest.mark.parametrize('name', ['const', 'jump', 'rewind'])
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_solver_reproduces_reference_run(golden_shifts, name, stype):
    g = golden_shifts
    src_pcm, dst_pcm, ev, params = scenario_inputs(g, name)
    src = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, stype)
    dst = TracingStream(oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, stype), src)
    events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(ev)]
    groups = prepare_search_groups(events, src.duration_seconds, [], 0.417, 0.417)
    want_groups = g['{0}_{1}_groups'.format(name, stype)]
    assert [[grp[0].source_index, grp[-1].source_index] for grp in groups] == want_groups.tolist()
    calculate_shifts(src, dst, groups, float(params[0]), float(params[1]), int(params[2]))
    # the same sequence of matcher calls, with the same arguments and results ...
    want_calls = g['{0}_{1}_calls'.format(name, stype)]
    got_calls = np.array(dst.calls, np.float64)
    assert got_calls.shape == want_calls.shape
    assert np.array_equal(got_calls[:, :4], want_calls[:, :4])
    assert np.abs(got_calls[:, 4] - want_calls[:, 4]).max() <= 5e-6
    assert np.abs(got_calls[:, 5] - want_calls[:, 5]).max() <= 1.0 / 12000 + 1e-12
    # ... and the same per-event outcome
    want = g['{0}_{1}_result'.format(name, stype)]
    got = np.array([[e.shift, e.diff, (e._link.source_index if e.linked else -1)] for e in events])
    assert np.array_equal(got[:, 2], want[:, 2])
    assert np.abs(got[:, 0] - want[:, 0]).max() <= 1.0 / 12000 + 1e-12
    assert np.abs(got[:, 1] - want[:, 1]).max() <= 5e-6


def test_prepare_search_groups_links_comments_duplicates_and_nested():
    ev = [ScriptEvent(0, 1.0, 3.0), ScriptEvent(1, 1.0, 3.0), ScriptEvent(2, 1.5, 2.5, is_comment=True),
          ScriptEvent(3, 4.0, 4.0), ScriptEvent(4, 5.0, 9.0), ScriptEvent(5, 6.0, 8.0), ScriptEvent(6, 50.0, 60.0),
          ScriptEvent(7, 9.1, 9.3), ScriptEvent(8, 9.35, 9.5), ScriptEvent(9, 12.0, 14.0)]
    groups = prepare_search_groups(ev, 40.0, [], 0.417, 0.417)
    assert ev[1]._link is ev[0]              # same start and end -> linked to the first (sushi.py:373-378)
    assert ev[2]._link is ev[3]              # comment -> next event (sushi.py:355-360)
    assert ev[3]._link is ev[4]              # zero duration -> next event (sushi.py:365-371)
    assert ev[6]._link is ev[5] or ev[6].linked   # beyond the audio -> last unlinked (sushi.py:361-364)
    assert ev[5]._link is ev[4]              # nested inside an earlier group (sushi.py:386-396)
    assert [[e.source_index for e in grp] for grp in groups] == [[0], [4], [7, 8], [9]]   # short lines merged


def test_nested_group_linking_fast_path_equals_reference_scan():
    """The monotonic-stack search for the enclosing group (sorted scripts) gives the same links as the
    reference's backwards scan (sushi.py:386-396), including ties and chains of nesting."""
    rng = np.random.default_rng(12)
    for trial in range(30):
        starts = np.sort(np.round(rng.uniform(0, 60, 80), 2))
        ends = starts + np.round(rng.choice([0.5, 1.0, 3.0, 8.0, 20.0], 80), 2)
        a = [ScriptEvent(i, float(s), float(e)) for i, (s, e) in enumerate(zip(starts, ends))]
        b = [ScriptEvent(i, float(s), float(e)) for i, (s, e) in enumerate(zip(starts, ends))]
        got = prepare_search_groups(a, 1000.0, [], 0.417, 0.417)
        # the reference's literal scan on the copy
        from sushi_b200.grouping import merge_short_lines_into_groups
        last = None
        for idx, e in enumerate(b):
            twin = next((x for x in reversed(b[:idx]) if x.start == e.start and not x.linked and x.end == e.end), None) \
                if idx and b[idx - 1].start == e.start else None
            if twin:
                e.link_event(twin)
        groups = merge_short_lines_into_groups([e for e in b if not e.linked], [], 0.417, 0.417)
        want = []
        for idx, g in enumerate(groups):
            outer = next((x for x in reversed(groups[:idx]) if x[0].start <= g[0].start and x[-1].end >= g[-1].end), None)
            if outer is None:
                want.append(g)
            else:
                for e in g:
                    e.link_event(outer[0])
        assert [[e.source_index for e in g] for g in got] == [[e.source_index for e in g] for g in want]
        assert [(e._link.source_index if e.linked else -1) for e in a] == [(e._link.source_index if e.linked else -1) for e in b]
