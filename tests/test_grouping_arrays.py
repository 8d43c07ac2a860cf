"""the column-wise heuristics of sushi_b200/grouping.py against the reference's
event-by-event loops (tests/list_heuristics.py) on a 10 000-event synthetic script: identical groups,
identical links, bit-identical averages -- for matcher-typed results (np.float32 diffs, float shifts),
plain Python floats, and mixed types (which take the scalar fallback)."""
import copy

import numpy as np
import pytest

from sushi_b200 import grouping
from tests import list_heuristics as ref


class Ev(object):
    __slots__ = ('start', 'end', 'shift', 'diff', 'linked')

    def __init__(self, start, end, shift, diff):
        self.start, self.end, self.shift, self.diff, self.linked = start, end, shift, diff, None

    def set_shift(self, shift, diff):
        self.shift, self.diff = shift, diff

    def link_event(self, other):
        self.linked = other


def make_script(count, seed, diff_type, broken_head=7, broken_tail=4, all_broken=False):
    """Piecewise-constant shifts with jitter below and jumps above ALLOWED_ERROR, including differences
    that sit within a few ulps of the threshold; diffs around 0.03 with outliers at both borders."""
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.uniform(0.5, 3.0, count))
    base = np.cumsum(np.where(rng.random(count) < 0.02, rng.choice([-3.0, 0.5, 0.02, 0.0100001, 0.01], count), 0.0))
    shift = base + rng.uniform(-0.004, 0.004, count)
    for i in range(50, count, 97):                     # exact threshold cases: next = this + 0.01 (+- 1 ulp)
        shift[i] = shift[i - 1] + 0.01
        if i + 1 < count:
            shift[i + 1] = np.nextafter(shift[i] + 0.01, np.inf)
    diff = rng.uniform(0.02, 0.05, count)
    if all_broken:
        diff[:] = rng.uniform(0.9, 1.0, count) * np.where(rng.random(count) < 0.5, 1.0, 0.001)
        diff[count // 2] = 0.0
    else:
        diff[:broken_head] = rng.uniform(0.5, 1.0, broken_head)
        diff[count - broken_tail:] = rng.uniform(0.0, 0.003, broken_tail)
    events = []
    for i in range(count):
        d = diff_type(diff[i]) if diff_type is not None else (np.float32(diff[i]) if i % 2 else float(diff[i]))
        events.append(Ev(float(t[i]), float(t[i] + 1.0), float(shift[i]), d))
    return events


def link_indices(events):
    pos = {id(e): i for i, e in enumerate(events)}
    return [None if e.linked is None else pos[id(e.linked)] for e in events]


TYPES = [np.float32, float, np.float64, None]


@pytest.mark.parametrize('diff_type', TYPES)
def test_detect_groups_10k(diff_type):
    events = make_script(10000, 1, diff_type)
    want = ref.detect_groups(events)
    got = grouping.detect_groups(events)
    assert [[id(e) for e in g] for g in got] == [[id(e) for e in g] for g in want]
    assert len(want) > 100
    # a generator works too (split_broken_groups hands lists, the reference accepts any iterable)
    got = grouping.detect_groups(e for e in events)
    assert [len(g) for g in got] == [len(g) for g in want]


def test_detect_groups_float32_shifts():
    events = make_script(2000, 5, np.float32)
    for e in events:
        e.shift = np.float32(e.shift)
    assert [len(g) for g in grouping.detect_groups(events)] == [len(g) for g in ref.detect_groups(events)]


def test_detect_groups_empty_raises_like_next():
    with pytest.raises(StopIteration):
        grouping.detect_groups([])
    with pytest.raises(StopIteration):
        ref.detect_groups([])


@pytest.mark.parametrize('diff_type', TYPES)
@pytest.mark.parametrize('all_broken', [False, True])
def test_fix_near_borders_10k(diff_type, all_broken):
    a = make_script(10000, 2, diff_type, all_broken=all_broken)
    b = copy.deepcopy(a)
    with np.errstate(all='ignore'):
        ref.fix_near_borders(a)
        grouping.fix_near_borders(b)
    assert link_indices(a) == link_indices(b)
    if not all_broken:
        assert link_indices(a)[:8] == [7] * 7 + [None] and link_indices(a)[-5:] == [None] + [10000 - 5] * 4


@pytest.mark.parametrize('diff_type', TYPES)
def test_average_shifts_10k_bit_identical(diff_type):
    a = make_script(10000, 3, diff_type)
    for i in range(0, 10000, 13):
        a[i].linked = a[i - 1] if i else a[1]
    b = copy.deepcopy(a)
    want = ref.average_shifts(a)
    got = grouping.average_shifts(b)
    assert type(got) is type(want) and np.float64(got).tobytes() == np.float64(want).tobytes()
    assert [e.shift for e in a] == [e.shift for e in b]


def test_whole_post_processing_chain_10k():
    """fix_near_borders -> smooth_events -> detect_groups -> average_shifts per group, the order of
    sushi.py:682-711 without chapters: same final shifts as the loops, event by event."""
    a = make_script(10000, 4, np.float32)
    b = copy.deepcopy(a)

    def chain(events, mod, smooth):
        mod.fix_near_borders(events)
        live = [e for e in events if not e.linked]
        smooth(live, 3)
        groups = mod.detect_groups(live)
        for g in groups:
            mod.average_shifts(g)
        return [len(g) for g in groups]

    ga = chain(a, ref, grouping.smooth_events)
    gb = chain(b, grouping, grouping.smooth_events)
    assert ga == gb
    assert [np.float64(e.shift).tobytes() for e in a] == [np.float64(e.shift).tobytes() for e in b]
    assert link_indices(a) == link_indices(b)
