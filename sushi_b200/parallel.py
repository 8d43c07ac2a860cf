"""Multi-GPU matching: one process per GPU, events sharded across ranks.

Every query is independent given (template offset, length, first lag, lag count), so each rank takes a
contiguous slice of the (time-sorted) event list -- neighbouring events search overlapping parts of the
destination stream, which keeps a rank's block spectra hot in its L2 and lets it transform only the part of
the stream its own events' windows cover.  Two collectives, both outside the kernels:

    broadcast   the normalised streams from the root rank           (65 MB uint8 per 90-minute stream)
    all_gather  the per-event (idx, diff) results                    (12 B per event)

There is no collective on the data path of a query, hence no fused compute + communication kernel.

`ShardedMatcher` is the orchestration (what is broadcast when, who matches what, how the padded gather is
unpacked); it is written against two small interfaces so that the SAME code runs
  * on GPUs:  `NcclComm` (the library's own NCCL communicator, sb_comm_* in include/sushi_b200.h -- no Python
    framework on the timed path) + `DeviceBackend` (WavStream / sb_find_batch_device), and
  * in the CPU test-suite: `TorchComm` over gloo + a NumPy backend supplied by the test
    (tests/test_parallel_cpu.py), which checks sharding, padding, ordering and the header exchange.
"""
import ctypes
import os
import time

import numpy as np

from ._nvtx import nvtx_range
from .common import SushiError
from .wavstream import StreamGeometry


def shard_bounds(count, world_size, rank):
    """Contiguous, balanced [lo, hi) slice of `count` items for `rank`; the first count % world
    ranks get one extra item."""
    base, extra = divmod(int(count), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(count, world_size):
    return [shard_bounds(count, world_size, r)[1] - shard_bounds(count, world_size, r)[0] for r in range(world_size)]


# ------------------------------------------------------------------------------------------------
# communicators
# ------------------------------------------------------------------------------------------------
def _rendezvous_path():
    """One file per job on this node: the ranks of a torchrun job share MASTER_PORT and their parent (the
    agent process); a later job on the same port has another parent or run id."""
    tag = '%s_%s_%s_%s' % (os.environ.get('MASTER_ADDR', 'local'), os.environ.get('MASTER_PORT', '0'),
                           os.environ.get('TORCHELASTIC_RUN_ID', 'none'), os.getppid())
    return os.path.join(os.environ.get('SUSHI_B200_RDZV_DIR', '/tmp'), 'sushi_b200_nccl_' + tag.replace('/', '_'))


class NcclComm(object):
    """The library's NCCL communicator (one process per GPU on one node).  The 128-byte NCCL id travels
    from rank 0 to the others through a file in /tmp -- the ranks of a single-node job share a file system,
    and that keeps PyTorch (or any other framework) out of the process."""

    def __init__(self, rank=None, world_size=None, lib=None, timeout_s=120.0):
        from . import _native
        self.rank = int(os.environ.get('RANK', '0')) if rank is None else int(rank)
        self.world_size = int(os.environ.get('WORLD_SIZE', '1')) if world_size is None else int(world_size)
        self._native = _native
        self.lib = lib if lib is not None else _native.lib()
        uid = (ctypes.c_uint8 * 128)()
        path = _rendezvous_path()
        if self.rank == 0:
            _native.check(self.lib.sb_comm_unique_id(uid), 'sb_comm_unique_id')
            tmp = path + '.tmp%d' % os.getpid()
            with open(tmp, 'wb') as f:
                f.write(bytes(uid))
            os.replace(tmp, path)
        else:
            deadline = time.time() + timeout_s
            while True:
                try:
                    with open(path, 'rb') as f:
                        raw = f.read()
                    if len(raw) == 128:
                        break
                except OSError:
                    pass
                if time.time() > deadline:
                    raise SushiError('NcclComm: rank %d never saw the NCCL id of rank 0 at %s' % (self.rank, path))
                time.sleep(0.01)
            ctypes.memmove(uid, raw, 128)
        _native.check(self.lib.sb_comm_init(uid, self.world_size, self.rank), 'sb_comm_init')
        self.barrier()
        if self.rank == 0:
            try:
                os.remove(path)
            except OSError:
                pass

    def broadcast(self, buf, nbytes, root, slot):
        self._native.check(self.lib.sb_comm_broadcast(ctypes.c_void_p(int(buf)), int(nbytes), int(root), int(slot)), 'sb_comm_broadcast')

    def wait(self, slot):
        self._native.check(self.lib.sb_comm_wait(int(slot)), 'sb_comm_wait')

    def all_gather(self, send, recv, nbytes):
        self._native.check(self.lib.sb_comm_all_gather(ctypes.c_void_p(int(send)), ctypes.c_void_p(int(recv)), int(nbytes)), 'sb_comm_all_gather')

    def max_over_ranks(self, values):
        a = (ctypes.c_float * len(values))(*[float(v) for v in values])
        self._native.check(self.lib.sb_comm_max_f32(a, len(values)), 'sb_comm_max_f32')
        return [float(v) for v in a]

    def barrier(self):
        self._native.check(self.lib.sb_comm_barrier(), 'sb_comm_barrier')

    def close(self):
        self.lib.sb_comm_destroy()


class TorchComm(object):
    """The same five operations over torch.distributed on NumPy-backed host buffers (gloo): what the CPU
    test-suite runs the orchestration on.  Buffers are NumPy uint8 arrays."""

    def __init__(self, dist, torch):
        self.dist, self.torch = dist, torch
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()

    def broadcast(self, buf, nbytes, root, slot):
        self.dist.broadcast(self.torch.from_numpy(buf[:nbytes]), root)

    def wait(self, slot):
        pass

    def all_gather(self, send, recv, nbytes):
        self.dist.all_gather_into_tensor(self.torch.from_numpy(recv[:nbytes * self.world_size]), self.torch.from_numpy(send[:nbytes]))

    def max_over_ranks(self, values):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def barrier(self):
        self.dist.barrier()

    def close(self):
        pass


class SingleComm(object):
    """world_size 1: nothing to exchange (the single-GPU bench runs the same ShardedMatcher)."""
    rank, world_size = 0, 1

    def broadcast(self, buf, nbytes, root, slot): pass
    def wait(self, slot): pass
    def max_over_ranks(self, values): return [float(v) for v in values]
    def barrier(self): pass
    def close(self): pass

    def __init__(self, backend=None):
        self._backend = backend

    def all_gather(self, send, recv, nbytes):
        self._backend.copy(recv, send, nbytes)


# ------------------------------------------------------------------------------------------------
# device backend
# ------------------------------------------------------------------------------------------------
class DeviceBackend(object):
    """Buffers in HBM, matching through WavStream (sb_stream_create_device + sb_find_batch_device)."""

    def __init__(self, lib=None):
        from . import _native
        from .wavstream import WavStream
        self._native, self._WavStream = _native, WavStream
        self.lib = lib if lib is not None else _native.lib()
        self._owned = []

    def alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._native.check(self.lib.sb_device_alloc(max(int(nbytes), 16), ctypes.byref(p)), 'sb_device_alloc')
        self._owned.append(p.value)
        return p.value

    def offset(self, buf, nbytes):
        return buf + int(nbytes)

    def upload(self, buf, host_array):
        a = np.ascontiguousarray(host_array)
        self._native.check(self.lib.sb_copy_to_device(ctypes.c_void_p(buf), a.ctypes.data_as(ctypes.c_void_p), a.nbytes), 'sb_copy_to_device')

    def download(self, buf, nbytes):
        out = np.empty(int(nbytes), np.uint8)
        self._native.check(self.lib.sb_copy_to_host(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(buf), int(nbytes)), 'sb_copy_to_host')
        return out

    def copy(self, dst, src, nbytes):
        self._native.check(self.lib.sb_copy_on_device(ctypes.c_void_p(dst), ctypes.c_void_p(src), int(nbytes)), 'sb_copy_on_device')

    def open_stream(self, buf, geom, sample_type):
        return self._WavStream.from_device(buf, geom.total_samples, sample_type, geom.sample_rate, geom.padding_size, geom.sample_count)

    def match(self, dst, src, plan, idx_buf, diff_buf):
        toff, tlen, lag0, nlags = plan
        dst.find_planned_device(src, toff, tlen, lag0, nlags, diff_buf, idx_buf)

    def close_stream(self, s):
        s.close()

    def release(self):
        for p in self._owned:
            self.lib.sb_device_free(ctypes.c_void_p(p))
        self._owned = []


# ------------------------------------------------------------------------------------------------
# orchestration
# ------------------------------------------------------------------------------------------------
_HEADER_WORDS = 12      # per stream: total, padding, rate, sample_count (as float64 bits), dtype code; + event count


class ShardedMatcher(object):
    """find_substream for a whole event list, sharded over the ranks of `comm`.

        m = ShardedMatcher(comm, backend)
        m.set_streams(src, dst)                 # root: objects with .data/.sample_rate/.padding_size/.sample_count
        diffs, times = m.find_batch(starts, ends, centers, windows)     # root passes arrays, others None

    set_streams puts the root's normalised streams into a backend buffer on every rank (header exchange +
    upload on the root); find_batch = broadcast of the streams, each rank's contiguous shard of the queries
    against them, all-gather of (idx, diff).  Every rank returns the full result arrays, in event order, equal to
    what the root's single-GPU find_substream_batch returns (a value depends only on template and position).
    The stream broadcast sits inside find_batch on purpose: it is part of the job the benchmark times."""

    def __init__(self, comm, backend, root=0):
        self.comm, self.backend, self.root = comm, backend, root
        self.geom = [None, None]
        self.sample_type = None
        self._bufs = [None, None]
        self._nbytes = [0, 0]
        self._hdr = backend.alloc(8 * _HEADER_WORDS)
        self._res = None
        self._res_cap = 0
        self._ev = None
        self._ev_cap = 0
        self.last_plan = None

    # -- small fixed-size exchanges go through the same buffers / collectives as the data -----------
    def _bcast_words(self, words):
        """np.float64[_HEADER_WORDS] from the root to everyone."""
        n = 8 * _HEADER_WORDS
        if self.comm.rank == self.root:
            self.backend.upload(self._hdr, np.asarray(words, np.float64))
        self.comm.broadcast(self._hdr, n, self.root, 3)
        self.comm.wait(3)
        return self.backend.download(self._hdr, n).view(np.float64).copy()

    def set_streams(self, src=None, dst=None):
        words = np.zeros(_HEADER_WORDS, np.float64)
        if self.comm.rank == self.root:
            for i, s in enumerate((src, dst)):
                words[5 * i:5 * i + 5] = [s.data.shape[1], s.padding_size, s.sample_rate, s.sample_count,
                                          0 if s.data.dtype == np.uint8 else 1]
        words = self._bcast_words(words)
        for i in range(2):
            total, pad, rate, count, code = words[5 * i:5 * i + 5]
            self.geom[i] = StreamGeometry(int(rate) if float(rate).is_integer() else rate, int(pad), count, int(total))
            stype = 'uint8' if code == 0 else 'float32'
            if i and stype != self.sample_type:
                raise SushiError('ShardedMatcher: source and destination streams differ in sample type')
            self.sample_type = stype
            nbytes = int(total) * (1 if code == 0 else 4)
            if self._bufs[i] is None or self._nbytes[i] != nbytes:
                self._bufs[i], self._nbytes[i] = self.backend.alloc(nbytes), nbytes
        if self.comm.rank == self.root:
            self.backend.upload(self._bufs[0], src.data)
            self.backend.upload(self._bufs[1], dst.data)

    def upload_streams(self, src, dst):
        """Root only: fresh host data for the same geometry (the end-to-end leg copies every step)."""
        if self.comm.rank == self.root:
            self.backend.upload(self._bufs[0], src.data)
            self.backend.upload(self._bufs[1], dst.data)

    def _share_events(self, starts, ends, centers, windows):
        words = np.zeros(_HEADER_WORDS, np.float64)
        if self.comm.rank == self.root:
            words[0] = len(starts)
        count = int(self._bcast_words(words)[0])
        nbytes = 4 * 8 * count
        if self._ev_cap < nbytes:
            self._ev, self._ev_cap = self.backend.alloc(nbytes), nbytes
        if count == 0:
            return [np.zeros(0)] * 4
        if self.comm.rank == self.root:
            self.backend.upload(self._ev, np.concatenate([np.asarray(a, np.float64) for a in (starts, ends, centers, windows)]))
        self.comm.broadcast(self._ev, nbytes, self.root, 2)
        self.comm.wait(2)
        ev = self.backend.download(self._ev, nbytes).view(np.float64)
        return [ev[i * count:(i + 1) * count].copy() for i in range(4)]

    def plan(self, starts=None, ends=None, centers=None, windows=None):
        """Share the root's event list, plan ALL queries on every rank (integer arithmetic on four numbers per
        stream: cheap, and every rank then knows every query's start time), remember this rank's shard."""
        starts, ends, centers, windows = self._share_events(starts, ends, centers, windows)
        src_g, dst_g = self.geom
        toff, tlen, lag0, nlags, t0 = dst_g.plan_queries(src_g, starts, ends, centers, windows) if len(starts) else \
            (np.zeros(0, np.int64),) * 4 + (np.zeros(0),)
        count = len(toff)
        lo, hi = shard_bounds(count, self.comm.world_size, self.comm.rank)
        cap = max(shard_sizes(count, self.comm.world_size)) if count else 0
        need = 12 * cap * (self.comm.world_size + 1)
        if self._res_cap < need:
            self._res, self._res_cap = self.backend.alloc(need), need
        self.last_plan = dict(count=count, lo=lo, hi=hi, cap=cap, t0=t0,
                              shard=tuple(np.ascontiguousarray(a[lo:hi]) for a in (toff, tlen, lag0, nlags)),
                              all=(toff, tlen, lag0, nlags))
        return self.last_plan

    def open_resident(self):
        """Broadcast both streams once and keep them open on every rank: later run_planned() calls match against
        the resident streams (running sums and block spectra are built once).  This is how the sequential solver
        and the window sweep (BASELINE config 5) issue many batches against one pair of streams."""
        self.close_resident()
        comm, be = self.comm, self.backend
        comm.broadcast(self._bufs[0], self._nbytes[0], self.root, 0)
        comm.broadcast(self._bufs[1], self._nbytes[1], self.root, 1)
        comm.wait(0)
        src = be.open_stream(self._bufs[0], self.geom[0], self.sample_type)
        comm.wait(1)
        dst = be.open_stream(self._bufs[1], self.geom[1], self.sample_type)
        self._resident = (src, dst)

    def close_resident(self):
        res = getattr(self, '_resident', None)
        if res is not None:
            self.backend.close_stream(res[0])
            self.backend.close_stream(res[1])
        self._resident = None

    def run_planned(self):
        """The timed part on device buffers: broadcast both streams, match this rank's shard, all-gather (with
        resident streams, open_resident(): the last two only).  Results stay in the backend buffer (rank-major,
        padded); read them with gather_results()."""
        p = self.last_plan
        comm, be = self.comm, self.backend
        resident = getattr(self, '_resident', None)
        if resident is not None:
            src, dst = resident
        else:
            with nvtx_range('sushi_b200: broadcast + open streams'):
                comm.broadcast(self._bufs[0], self._nbytes[0], self.root, 0)
                comm.broadcast(self._bufs[1], self._nbytes[1], self.root, 1)       # overlaps the source stream's running sums
                comm.wait(0)
                src = be.open_stream(self._bufs[0], self.geom[0], self.sample_type)
                comm.wait(1)
                dst = be.open_stream(self._bufs[1], self.geom[1], self.sample_type)
        cap = p['cap']
        send = self._res                                                   # [cap x int64 idx][cap x float32 diff]
        if p['hi'] > p['lo']:
            with nvtx_range('sushi_b200: match shard'):
                be.match(dst, src, p['shard'], send, be.offset(send, 8 * cap))
        if cap:
            with nvtx_range('sushi_b200: all-gather'):
                comm.all_gather(send, be.offset(self._res, 12 * cap), 12 * cap)
        if resident is None:
            be.close_stream(src)
            be.close_stream(dst)

    def gather_results(self):
        p = self.last_plan
        count, cap, world = p['count'], p['cap'], self.comm.world_size
        if count == 0:
            return np.zeros(0, np.float32), np.zeros(0, np.int64)
        raw = self.backend.download(self.backend.offset(self._res, 12 * cap), 12 * cap * world)
        diff = np.empty(count, np.float32)
        idx = np.empty(count, np.int64)
        for r in range(world):
            lo, hi = shard_bounds(count, world, r)
            blk = raw[12 * cap * r:12 * cap * (r + 1)]
            idx[lo:hi] = blk[:8 * cap].view(np.int64)[:hi - lo]
            diff[lo:hi] = blk[8 * cap:].view(np.float32)[:hi - lo]
        return diff, idx

    def find_batch(self, starts=None, ends=None, centers=None, windows=None):
        """(diffs float32[count], times float64[count]) on every rank -- WavStream.find_substream_batch, sharded."""
        self.plan(starts, ends, centers, windows)
        self.run_planned()
        diff, idx = self.gather_results()
        return diff, self.last_plan['t0'] + idx / float(self.geom[1].sample_rate)
