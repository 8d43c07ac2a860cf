#!/usr/bin/env python3
"""Small fixed workload for profiler captures: config-2 shaped streams (10 min instead of 30 to
keep set-up short), `--events` queries at +-60 s, a few batches through the public batched API.
    ncu --set full --import-source on -k regex:k_match_fused -s 1 -c 1 -o gpurun_out/fused python tools/ncu_case.py
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sushi_b200 import WavStream, synth, _native   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--events', type=int, default=300)
ap.add_argument('--duration', type=float, default=600.0)
ap.add_argument('--window', type=float, default=60.0)
ap.add_argument('--batches', type=int, default=3)
ap.add_argument('--engine', type=int, default=2)
ap.add_argument('--block', type=int, default=16384)
ap.add_argument('--sample-type', default='uint8')
ap.add_argument('--hop-mode', type=int, default=1)
ap.add_argument('--premac-mode', type=int, default=0)
ap.add_argument('--min-len', type=float, default=1.0)
ap.add_argument('--max-len', type=float, default=4.0)
a = ap.parse_args()

src_pcm, dst_pcm = synth.make_pair(a.duration, 2, 1.5)
src = WavStream.from_pcm(src_pcm, 12000, sample_type=a.sample_type)
dst = WavStream.from_pcm(dst_pcm, 12000, sample_type=a.sample_type)
lib = _native.lib()
_native.check(lib.sb_set_block_size(a.block))
_native.check(lib.sb_set_engine(a.engine))
_native.check(lib.sb_set_hop_mode(a.hop_mode))
_native.check(lib.sb_set_premac_mode(a.premac_mode))
starts, ends = synth.make_events(a.events, a.duration, 2, a.min_len, a.max_len)
for _ in range(a.batches):
    d, t = dst.find_substream_batch(src, starts, ends, starts, np.full(len(starts), a.window))
ok = (ends + 1.5 < a.duration)
print('events', len(starts), 'max |shift-1.5| (samples):', float(np.abs((t - starts)[ok] - 1.5).max() * 12000))
