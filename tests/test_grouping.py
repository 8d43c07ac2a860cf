"""The grouping heuristics (kept intact around the matcher) against the reference's own unit tests
(/root/reference/tests/main.py:34-165, ported to Python 3 / pytest; same inputs, same expectations)."""
import pytest

from sushi_b200 import grouping
from sushi_b200.common import SushiError


class FakeEvent(object):           # tests/main.py:12-31
    def __init__(self, shift=0.0, diff=0.0, end=0.0, start=0.0):
        self.shift, self.linked, self.diff, self.start, self.end = shift, None, diff, start, end

    def set_shift(self, shift, diff):
        self.shift, self.diff = shift, diff

    def link_event(self, other):
        self.linked = other

    def __repr__(self):
        return repr(self.shift)

    def __eq__(self, other):
        return self.__dict__ == other.__dict__

    __hash__ = object.__hash__


def same_items(a, b):
    a, b = list(a), list(b)
    return len(a) == len(b) and all(x in b for x in a) and all(x in a for x in b)


# interpolate_nones -- tests/main.py:34-57
def test_interpolate_empty():
    assert grouping.interpolate_nones([], []) == []


def test_interpolate_no_valid_points():
    assert not grouping.interpolate_nones([None, None, None], [1, 2, 3])


def test_interpolate_no_nones():
    assert grouping.interpolate_nones([1, 2, 3], [1, 2, 3]) == [1, 2, 3]


@pytest.mark.parametrize('data,points,want', [
    ([1, None, 3, None, 5], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5]),
    ([1, None, None, None, 5], [1, 2, 3, 4, 5], [1, 2, 3, 4, 5]),
    ([None, None, 2, None, None], [1, 2, 3, 4, 5], [2, 2, 2, 2, 2]),
    ([None, 0, 0, 0, None], [1, 2, 3, 4, 5], [0, 0, 0, 0, 0]),
    ([1, None, 10], [1, 2, 10], [1, 2, 10]),
])
def test_interpolate_cases(data, points, want):
    assert grouping.interpolate_nones(data, points) == want


# running_median / smooth_events -- tests/main.py:60-82
def test_running_median_keeps_borders():
    shifts = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    assert grouping.running_median(shifts, 5) == shifts


def test_running_median_removes_outliers():
    assert grouping.running_median([0.1, 0.1, 0.1, 9001, 0.1, 0.1, 0.1], 5) == [0.1] * 7


def test_running_median_needs_odd_window():
    with pytest.raises(SushiError):
        grouping.running_median([1, 2, 3], 4)


def test_smooth_events():
    events = [FakeEvent(x, diff=x) for x in (0.1, 0.1, 0.1, 9001, 7777, 0.1, 0.1, 0.1)]
    diffs = [e.diff for e in events]
    grouping.smooth_events(events, 7)
    assert [e.shift for e in events] == [0.1] * 8
    assert [e.diff for e in events] == diffs


# detect_groups -- tests/main.py:85-96
def test_detect_groups():
    events = [FakeEvent(0.5)] * 3 + [FakeEvent(1.0)] * 10 + [FakeEvent(0.5)] * 5
    assert [len(g) for g in grouping.detect_groups(events)] == [3, 10, 5]
    assert [len(g) for g in grouping.detect_groups([FakeEvent(0.5)] * 10)] == [10]


# groups_from_chapters -- tests/main.py:99-120
def test_groups_from_chapters():
    events = [FakeEvent(end=1), FakeEvent(end=2), FakeEvent(end=3)]
    groups = grouping.groups_from_chapters(events, [])
    assert len(groups) == 1 and groups[0] == events
    groups = grouping.groups_from_chapters(events, [0.0, 1.5])
    assert len(groups) == 2 and same_items([events[0]], groups[0]) and same_items(events[1:], groups[1])
    events = [FakeEvent(end=x) for x in range(1, 10)]
    groups = grouping.groups_from_chapters(events, [0.0, 3.2, 4.4, 7.7])
    assert len(groups) == 4
    assert same_items(events[0:3], groups[0]) and same_items(events[3:4], groups[1])
    assert same_items(events[4:7], groups[2]) and same_items(events[7:9], groups[3])


# split_broken_groups -- tests/main.py:123-151
def test_split_broken_groups():
    groups = [[FakeEvent(0.5), FakeEvent(0.5)], [FakeEvent(10.0)]]
    assert same_items(groups, grouping.split_broken_groups(groups))
    groups = [[FakeEvent(0.5)] * 10 + [FakeEvent(10.0)] * 5, [FakeEvent(0.5)] * 10]
    assert same_items([[FakeEvent(0.5)] * 10, [FakeEvent(10.0)] * 5, [FakeEvent(0.5)] * 10],
                      grouping.split_broken_groups(groups))
    groups = [[FakeEvent(0.5), FakeEvent(10.0)], [FakeEvent(10.0), FakeEvent(10.0), FakeEvent(15.0)]]
    assert same_items([[FakeEvent(0.5)], [FakeEvent(10.0)] * 3, [FakeEvent(15.0)]],
                      grouping.split_broken_groups(groups))


# fix_near_borders -- tests/main.py:154-165
def test_fix_near_borders():
    events = [FakeEvent(diff=x) for x in (0.9, 0.9, 0.1, 0.1, 0.1, 0.1, 0.1, 1.0, 0.9)]
    grouping.fix_near_borders(events)
    sf, sl = events[2], events[-3]
    assert [e.linked for e in events] == [sf, sf, None, None, None, None, None, sl, sl]
    events = [FakeEvent(diff=x) for x in (0.9, 0.9, 0.9, 1.0, 0.9)]
    grouping.fix_near_borders(events)
    assert [e.linked for e in events] == [None] * 5


def test_average_shifts_is_weighted_by_one_minus_diff():
    a, b = FakeEvent(1.0, diff=0.5), FakeEvent(2.0, diff=0.0)
    a.linked = b.linked = None
    avg = grouping.average_shifts([a, b])
    assert abs(avg - (1.0 * 0.5 + 2.0 * 1.0) / 1.5) < 1e-12 and a.shift == b.shift == avg


def test_vectorised_running_median_equals_the_loop():
    import numpy as np
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 7, 8, 50, 501):
        vals = list(rng.normal(0, 1, n))
        for w in (1, 3, 7, 11):
            half = w // 2
            want = [np.median(vals[i - min(half, i, n - i - 1):i + min(half, i, n - i - 1) + 1]) for i in range(n)]
            assert grouping.running_median(vals, w) == want
