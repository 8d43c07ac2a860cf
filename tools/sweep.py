#!/usr/bin/env python3
"""BASELINE config 5: event length x search span sweep on one GPU (device-timed, streams resident).
Writes gpurun_out/sweep.json and prints a markdown table: events/s, algorithmic GB/s (SURVEY.md 8d
bytes) and fraction of the measured HBM peak per cell.
    python tools/sweep.py [--queries 256] [--duration 1800]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sushi_b200 import WavStream, synth, _native   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--queries', type=int, default=256)
ap.add_argument('--duration', type=float, default=1800.0)
ap.add_argument('--sample-type', default='uint8')
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--hop-mode', type=int, default=1)
ap.add_argument('--premac-mode', type=int, default=0)
ap.add_argument('--events', default='0.5,1,3,10,30')
ap.add_argument('--windows', default='5,10,30,60,120,300,600')
ap.add_argument('--engine', type=int, default=-1, help='library engine (default: the library default)')
ap.add_argument('--epilogue', type=int, default=0, help='screening loop of the packed kernels: 1 | 2 (default: the library default)')
ap.add_argument('--out', default='sweep.json', help='file name under gpurun_out/')
a = ap.parse_args()

peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except (OSError, ValueError, KeyError):
    pass
src_pcm, dst_pcm = synth.make_pair(a.duration, 2, 1.5)
src = WavStream.from_pcm(src_pcm, 12000, sample_type=a.sample_type)
dst = WavStream.from_pcm(dst_pcm, 12000, sample_type=a.sample_type)
lib = _native.lib()
if a.engine >= 0:
    _native.check(lib.sb_set_engine(a.engine))
if a.epilogue > 0:
    _native.check(lib.sb_set_epilogue(a.epilogue))
_native.check(lib.sb_set_hop_mode(a.hop_mode))
_native.check(lib.sb_set_premac_mode(a.premac_mode))
EV = [float(x) for x in a.events.split(',')]
WIN = [float(x) for x in a.windows.split(',')]
bps = 1 if a.sample_type == 'uint8' else 4
pd, pi = ctypes.c_void_p(), ctypes.c_void_p()
_native.check(lib.sb_device_alloc(4 * a.queries, ctypes.byref(pd)))
_native.check(lib.sb_device_alloc(8 * a.queries, ctypes.byref(pi)))
rng = np.random.default_rng(5)
rows = []
for ev_len in EV:
    for win in WIN:
        starts = np.sort(rng.uniform(win * 0.25, a.duration - ev_len - 2.0, a.queries))
        starts = np.round(starts * 100) / 100
        ends = starts + ev_len
        toff, tlen, lag0, nlags, t0 = dst.plan_queries(src, starts, ends, starts, np.full(a.queries, win))
        alg = float(np.sum(bps * tlen + bps * (nlags + tlen - 1) + 16))
        best = None
        for rep in range(a.reps + 1):
            _native.check(lib.sb_sync())
            _native.check(lib.sb_timer_start())
            dst.find_planned_device(src, toff, tlen, lag0, nlags, pd.value, pi.value)
            ms = ctypes.c_float()
            _native.check(lib.sb_timer_stop(ctypes.byref(ms)))
            if rep > 0:
                best = ms.value if best is None else min(best, ms.value)
        d, i = dst.find_planned(src, toff, tlen, lag0, nlags)
        shift = (t0 + i / 12000.0) - starts
        ok = (ends + 1.5 < a.duration) & (np.abs(1.5) <= win)
        good = float(np.mean(np.abs(shift[ok] - 1.5) <= 1.0 / 12000 + 1e-9)) if ok.any() else float('nan')
        rows.append({'event_s': ev_len, 'window_s': win, 'queries': a.queries, 'lags': int(np.median(nlags)),
                     'ms': round(best, 4), 'events_per_s': round(a.queries / best * 1e3, 1),
                     'alg_gbs': round(alg / best / 1e6, 2), 'frac_hbm': round(alg / best / 1e6 / peak, 5),
                     'shift_recovered': good})
        print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump({'peak_hbm_gbs': peak, 'sample_type': a.sample_type, 'duration_s': a.duration, 'engine': lib.sb_get_engine(), 'cells': rows},
          open(os.path.join(ROOT, 'gpurun_out', a.out), 'w'), indent=1)
print('\n| event \\ window | ' + ' | '.join('±%g s' % w for w in WIN) + ' |')
print('|---|' + '---|' * len(WIN))
for ev_len in EV:
    cells = [r for r in rows if r['event_s'] == ev_len]
    print('| %g s | ' % ev_len + ' | '.join('%.0f ev/s, %.0f GB/s (%.1f%%)' % (r['events_per_s'], r['alg_gbs'], 100 * r['frac_hbm']) for r in cells) + ' |')
