"""The N>1 host logic on CPU: gloo processes run parallel.ShardedMatcher -- the orchestration the GPU path uses --
with a NumPy backend whose matcher is the CPU oracle; the gathered result equals the single-process run."""
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


class OracleBackend(object):
    """Test stand-in for parallel.DeviceBackend: buffers are NumPy uint8 arrays on the host, the matcher is the
    CPU oracle (the reference's find_substream arithmetic over cv2).  Test infrastructure only."""

    def alloc(self, nbytes):
        return np.zeros(max(int(nbytes), 16), np.uint8)

    def offset(self, buf, nbytes):
        return buf[int(nbytes):]

    def upload(self, buf, host_array):
        raw = np.ascontiguousarray(host_array).reshape(-1).view(np.uint8)
        buf[:raw.size] = raw

    def download(self, buf, nbytes):
        return buf[:int(nbytes)].copy()

    def copy(self, dst, src, nbytes):
        dst[:nbytes] = src[:nbytes]

    def open_stream(self, buf, geom, sample_type):
        dt = np.uint8 if sample_type == 'uint8' else np.float32
        return buf[:geom.total_samples * np.dtype(dt).itemsize].view(dt)[None, :]

    def match(self, dst, src, plan, idx_buf, diff_buf):
        import cv2
        toff, tlen, lag0, nlags = plan
        n = len(toff)
        idx, diff = idx_buf[:8 * n].view(np.int64), diff_buf[:4 * n].view(np.float32)
        for q in range(n):
            cur = cv2.matchTemplate(dst[:, lag0[q]:lag0[q] + nlags[q] + tlen[q] - 1], src[:, toff[q]:toff[q] + tlen[q]],
                                    cv2.TM_SQDIFF_NORMED)[0]
            idx[q] = cur.argmin()
            diff[q] = cur[idx[q]]

    def close_stream(self, s):
        pass


def _worker(rank, world, port, out_path, count):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from sushi_b200 import synth
    from tests.helpers import oracle_stream_from_pcm
    dist.init_process_group('gloo', rank=rank, world_size=world)
    m = parallel.ShardedMatcher(parallel.TorchComm(dist, torch), OracleBackend())
    # rank 0 owns the streams and the event list; the others learn everything through the matcher
    if rank == 0:
        src_pcm, dst_pcm = synth.make_pair(16.0, 3, 0.75)
        rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, 'uint8')
        rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, 'uint8')
        m.set_streams(rs, rd)
        starts = np.array([1.0, 2.2, 4.0, 6.5, 8.0, 9.1, 11.0])[:count]
        ends = starts + np.array([1.5, 0.7, 2.0, 1.0, 0.9, 1.3, 2.5])[:count]
        diffs, times = m.find_batch(starts, ends, starts, np.full(len(starts), 2.0))
        want = [rd.find_substream(rs.get_substream(a, b), a, 2.0) for a, b in zip(starts, ends)]
        ok = all(float(d) == float(w[0]) and t == w[1] for d, t, w in zip(diffs, times, want)) and len(diffs) == count
        # the same list again against RESIDENT streams (one broadcast, then batches): identical answers
        m.open_resident()
        d2, t2 = m.find_batch(starts, ends, starts, np.full(len(starts), 2.0))
        d3, t3 = m.find_batch(starts[::-1].copy(), ends[::-1].copy(), starts[::-1].copy(), np.full(len(starts), 2.0))
        m.close_resident()
        ok = ok and np.array_equal(d2, diffs) and np.array_equal(t2, times) and np.array_equal(d3[::-1], diffs)
        np.savez(out_path, ok=np.array([ok]), shift=times - starts, lo_hi=np.array([m.last_plan['lo'], m.last_plan['hi']]))
    else:
        m.set_streams()
        diffs, times = m.find_batch()
        m.open_resident()
        d2, t2 = m.find_batch()
        m.find_batch()
        m.close_resident()
        assert np.array_equal(d2, diffs) and np.array_equal(t2, times)
        assert len(diffs) == count and m.geom[1].total_samples == 16 * 12000 + 240000 and m.sample_type == 'uint8'
        np.savez(out_path + '.rank%d.npz' % rank, diffs=diffs, times=times, lo_hi=np.array([m.last_plan['lo'], m.last_plan['hi']]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,count', [(2, 7), (3, 2), (2, 0)])
def test_gloo_sharded_matcher_equals_the_single_process_oracle(tmp_path, world, count):
    """ShardedMatcher (the code the GPU path runs, over gloo and a NumPy backend here): header and event
    exchange, contiguous shards (one rank may get none), padded all-gather, results in event order on every
    rank, identical to the reference's find_substream on rank 0's streams."""
    import torch.multiprocessing as mp
    port = _free_port()
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(world, port, out, count), nprocs=world, join=True)
    r = np.load(out)
    assert bool(r['ok'][0])
    if count:
        assert np.all(np.abs(r['shift'] - 0.75) <= 1.0 / 12000 + 1e-9)        # +0.75 s recovered by every shard
    other = np.load(out + '.rank1.npz')
    assert len(other['diffs']) == count
    assert tuple(r['lo_hi']) == parallel.shard_bounds(count, world, 0) and tuple(other['lo_hi']) == parallel.shard_bounds(count, world, 1)


def test_single_comm_runs_the_same_orchestration():
    from sushi_b200 import synth
    from tests.helpers import oracle_stream_from_pcm
    src_pcm, dst_pcm = synth.make_pair(12.0, 4, -0.5)
    rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, 'float32')
    rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, 'float32')
    be = OracleBackend()
    m = parallel.ShardedMatcher(parallel.SingleComm(be), be)
    m.set_streams(rs, rd)
    starts = np.array([2.0, 5.0, 7.5])
    diffs, times = m.find_batch(starts, starts + 1.2, starts, np.full(3, 3.0))
    for a, d, t in zip(starts, diffs, times):
        wd, wt = rd.find_substream(rs.get_substream(a, a + 1.2), a, 3.0)
        assert float(d) == float(wd) and t == wt
