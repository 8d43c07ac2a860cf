#!/usr/bin/env python3
"""Multi-GPU parity check of sushi_b200.parallel.ShardedMatcher over the library's NCCL communicator:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_sharded_gpu.py

(torchrun is only the launcher: it sets RANK / WORLD_SIZE / LOCAL_RANK; this process never imports torch.)
Rank 0 owns the streams and the event list; every rank must end up with the answers rank 0's single-GPU
find_substream_batch gives, bit for bit, and rank 0 checks a sample against the CPU oracle (cv2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sushi_b200 import WavStream, parallel, synth, _native     # noqa: E402

rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
lib = _native.lib(int(os.environ.get('LOCAL_RANK', '0')))
be = parallel.DeviceBackend(lib)
comm = parallel.NcclComm(rank, world, lib) if world > 1 else parallel.SingleComm(be)
m = parallel.ShardedMatcher(comm, be)
dur, nev, win = 600.0, 777, 45.0
ok = True
for stype in ('uint8', 'float32'):
    if rank == 0:
        from tests.helpers import oracle_stream_from_pcm
        src_pcm, dst_pcm = synth.make_pair(dur, 31, -2.25)
        rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, stype)
        rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, stype)
        starts, ends = synth.make_events(nev, dur, 32, 0.6, 5.0)
        m.set_streams(rs, rd)
        got = m.find_batch(starts, ends, starts, np.full(nev, win))
        src = WavStream.from_array(rs.data, 12000, rs.padding_size, rs.sample_count)
        dst = WavStream.from_array(rd.data, 12000, rd.padding_size, rd.sample_count)
        want = dst.find_substream_batch(src, starts, ends, starts, np.full(nev, win))
        same = np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        worst_d = worst_t = 0.0
        for q in np.linspace(0, nev - 1, 12).astype(int):
            d, t = rd.find_substream(rs.get_substream(starts[q], ends[q]), starts[q], win)
            worst_d, worst_t = max(worst_d, abs(float(d) - float(got[0][q]))), max(worst_t, abs(t - got[1][q]))
        good = same and worst_d <= 1e-5 and worst_t <= 1.0 / 12000 + 1e-9
        print('[rank 0] %s: %d events over %d rank(s): sharded == single-GPU call: %s; vs oracle max |ddiff| %.2e, max |dt| %.2e s -> %s'
              % (stype, nev, world, same, worst_d, worst_t, 'OK' if good else 'FAIL'), flush=True)
        ok = ok and good
        checksum = np.array([float(np.sum(got[0].astype(np.float64))), float(np.sum(got[1]))])
        src.close(); dst.close()
    else:
        m.set_streams()
        got = m.find_batch()
        checksum = np.array([float(np.sum(got[0].astype(np.float64))), float(np.sum(got[1]))])
    # every rank holds the same answers: the maximum over ranks of (+x, -x) pins equality
    a = comm.max_over_ranks([checksum[0] % 1e6, -(checksum[0] % 1e6), checksum[1] % 1e6, -(checksum[1] % 1e6)])
    same_everywhere = abs(a[0] + a[1]) < 1e-3 and abs(a[2] + a[3]) < 1e-3
    if rank == 0:
        print('[rank 0] %s: all %d ranks hold identical results: %s (NCCL %d)' % (stype, world, same_everywhere, lib.sb_comm_nccl_version()), flush=True)
    ok = ok and same_everywhere
comm.barrier()
comm.close()
sys.exit(0 if ok else 1)
