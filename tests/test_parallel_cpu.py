"""The N>1 host logic on CPU: world_size-2 gloo processes shard a planned query list, each runs the
CPU oracle on its shard, and the all-gathered result equals the single-process run."""
import os
import socket
import sys

import numpy as np
import pytest

from sushi_b200 import parallel


def test_shard_bounds_cover_everything_contiguously():
    for count in (0, 1, 7, 8, 9, 1250, 10000):
        for world in (1, 2, 3, 8):
            edges = [parallel.shard_bounds(count, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == count
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == parallel.shard_sizes(count, world)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from sushi_b200 import synth
    from tests.helpers import oracle_stream_from_pcm
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # rank 0 owns the streams; the others receive them by broadcast
    n = 16 * 12000 + 240000
    if rank == 0:
        src_pcm, dst_pcm = synth.make_pair(16.0, 3, 0.75)
        rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, 'uint8')
        rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, 'uint8')
        t_src, t_dst = torch.from_numpy(rs.data[0].copy()), torch.from_numpy(rd.data[0].copy())
        assert t_src.numel() == n
    else:
        t_src, t_dst = torch.empty(n, dtype=torch.uint8), torch.empty(n, dtype=torch.uint8)
    parallel.broadcast_stream(dist, t_src)
    parallel.broadcast_stream(dist, t_dst)
    from oracle.ref_matcher import RefStream
    src = RefStream(t_src.numpy()[None, :], 12000, 120000, 16 * 12000)
    dst = RefStream(t_dst.numpy()[None, :], 12000, 120000, 16 * 12000)
    starts = np.array([1.0, 2.2, 4.0, 6.5, 8.0, 9.1, 11.0])
    toff = np.array([src.sample_for_time(a) for a in starts])
    tlen = np.full(len(starts), 18000)
    lag0 = np.maximum(toff - 24000, 0)
    nlags = np.full(len(starts), 48001)

    def match(o, l, s, c):
        import cv2
        d = np.empty(len(o), np.float32)
        i = np.empty(len(o), np.int64)
        for q in range(len(o)):
            cur = cv2.matchTemplate(dst.data[:, s[q]:s[q] + c[q] + l[q] - 1], src.data[:, o[q]:o[q] + l[q]], cv2.TM_SQDIFF_NORMED)[0]
            i[q] = cur.argmin()
            d[q] = cur[i[q]]
        return d, i
    d, i = parallel.sharded_find(dist, torch, rank, world, match, toff, tlen, lag0, nlags)
    if rank == 0:
        d1, i1 = match(toff, tlen, lag0, nlags)
        np.savez(out_path, ok=np.array([np.array_equal(d.numpy(), d1) and np.array_equal(i.numpy(), i1)]),
                 shift=(i.numpy() + lag0 - toff))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_find(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    assert bool(r['ok'][0])
    assert np.all(np.abs(r['shift'] - 9000) <= 1)        # +0.75 s recovered by both shards
