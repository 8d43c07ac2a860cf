#!/usr/bin/env python3
"""BASELINE config 4 at full size on one GPU: 500 chapter groups x 20 events, per-chapter shift with a
slow drift, +-300 s window, through prepare_search_groups -> calculate_shifts -> chapter grouping and the
post-processing heuristics.  Checks the known per-chapter shifts and reports script events/s."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sushi_b200 import WavStream, synth, grouping, _hostmem   # noqa: E402
from sushi_b200.events import ScriptEvent                     # noqa: E402
from sushi_b200.grouping import prepare_search_groups         # noqa: E402
from sushi_b200.shifts import calculate_shifts                # noqa: E402

_hostmem.keep_heap()
n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 500
per_group = 20
dur = n_groups * 10.8
rng = np.random.default_rng(4)
chapters = [i * (dur / n_groups) for i in range(n_groups)]
# piecewise drift: shifts wander by <= 0.2 s per chapter, with a few jumps of several seconds
steps = rng.uniform(-0.2, 0.2, n_groups)
steps[rng.choice(n_groups, 12, replace=False)] += rng.uniform(-8, 8, 12)
shifts = np.round(np.cumsum(steps) * 12000) / 12000
shifts -= shifts.mean()
t0 = time.time()
src_pcm, dst_pcm = synth.make_pair(dur, 4, list(zip(chapters, shifts)))
src = WavStream.from_pcm(src_pcm, 12000)
dst = WavStream.from_pcm(dst_pcm, 12000)
starts, ends = synth.make_events(n_groups * per_group, dur, 4, 0.45, 0.9, 1.0)
events = [ScriptEvent(i, float(a), float(b)) for i, (a, b) in enumerate(zip(starts, ends))]
print('setup %.1f s, %d events, %d chapters, duration %.0f s' % (time.time() - t0, len(events), n_groups, dur))
t0 = time.time()
groups = prepare_search_groups(events, src.duration_seconds, chapters, 0.417, 0.417)
t1 = time.time()
calculate_shifts(src, dst, groups, 10.0, 300.0, 5)
t2 = time.time()
ev = [e for e in events if not e.linked]
by_chapter = grouping.groups_from_chapters(ev, chapters)
for grp in by_chapter:
    grouping.fix_near_borders(grp)
    grouping.smooth_events([e for e in grp if not e.linked], 3)
by_chapter = grouping.split_broken_groups(by_chapter)
for grp in by_chapter:
    grouping.average_shifts(grp)
t3 = time.time()
mid = (starts + ends) / 2
truth = shifts[np.searchsorted(chapters, mid, side='right') - 1]
got = np.array([e.shift for e in events])
inside = np.array([np.searchsorted(chapters, a, side='right') == np.searchsorted(chapters, b + 0.01, side='right')
                   for a, b in zip(starts, ends)])
err = np.abs(got - truth)
print('search groups %d; prepare %.2f s, calculate_shifts %.2f s (%.0f script events/s), heuristics %.2f s'
      % (len(groups), t1 - t0, t2 - t1, len(events) / (t2 - t1), t3 - t2))
print('events inside one chapter: %d; within 1 sample of the chapter shift: %.4f; within 10 ms: %.4f; max err %.4f s'
      % (inside.sum(), np.mean(err[inside] <= 1.0 / 12000 + 1e-9), np.mean(err[inside] <= 0.011), err[inside].max()))
