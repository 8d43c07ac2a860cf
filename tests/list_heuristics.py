"""Test infrastructure: the reference's event-by-event loops for three heuristics, restated for
Python 3 (sushi.py:120-127 detect_groups, :190-216 fix_near_borders, :309-316 average_shifts).
The product (sushi_b200/grouping.py) computes the same things from whole columns; the tests demand
identical groups, links and bit-identical averages.  Never imported by the product."""
import numpy as np

ALLOWED_ERROR = 0.01


def detect_groups(events_iter):                       # sushi.py:120-127
    events_iter = iter(events_iter)
    groups_list = [[next(events_iter)]]
    for event in events_iter:
        if abs(event.shift - groups_list[-1][-1].shift) > ALLOWED_ERROR:
            groups_list.append([])
        groups_list[-1].append(event)
    return groups_list


def fix_near_borders(events):                         # sushi.py:190-216
    def fix_border(event_list, median_diff):
        last_ten_diff = np.median([x.diff for x in event_list[:10]])
        diff_limit = min(last_ten_diff, median_diff)
        broken = []
        for event in event_list:
            if not 0.2 < (event.diff / diff_limit) < 5:
                broken.append(event)
            else:
                for x in broken:
                    x.link_event(event)
                return len(broken)
        return 0

    median_diff = np.median([x.diff for x in events])
    return fix_border(events, median_diff), fix_border(list(reversed(events)), median_diff)


def average_shifts(events):                           # sushi.py:309-316
    events = [e for e in events if not e.linked]
    shifts = [x.shift for x in events]
    weights = [1 - x.diff for x in events]
    avg = np.average(shifts, weights=weights)
    for e in events:
        e.set_shift(avg, e.diff)
    return avg
