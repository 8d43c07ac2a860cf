#!/usr/bin/env python3
"""Benchmark of the audio template-matching hot path (BASELINE.json metric: subtitle events/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload config2]

One "step" = one pass of the hot path over one batch of synthetic subtitle events: stream
preparation (running sums + block spectra) and every event's TM_SQDIFF_NORMED search at the
configuration's full window.  One event = one search group = one find_substream query
(SURVEY.md section 8d).  At N ranks every rank searches its own event list against the same two
streams (weak scaling: per-GPU work fixed), rank 0 broadcasts the streams over NCCL inside the
step and the per-event results are all-gathered.

Numbers printed (one JSON line, rank 0):
  value   events/s with the raw streams already resident in HBM, device-timed (CUDA events)
  e2e     events/s through the public Python API (WavStream.from_array + find_substream_batch)
          from page-locked HOST buffers: H2D of both streams and the query descriptors, D2H of
          the results, host-side planning -- all inside the timed region
  roofline       dominant kernel class against the measured HBM peak (algorithmic bytes, 8d)
  cpu_baseline   the oracle port (reference find_substream over cv2.matchTemplate) on this
                 box's host cores, on a bounded sample of the same events (N=1, rank 0 only)

--impl reference times that same CPU path as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sushi_b200 import _hostmem, parallel, synth      # noqa: E402
from sushi_b200.wavstream import WavStream   # noqa: E402

_hostmem.keep_heap()

SAMPLE_RATE = 12000
WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    'config1': dict(events=100, duration=60.0, window=10.0, min_len=1.0, max_len=4.0, shift=1.5, scaling='weak',
                    text='100 events, 2x60 s 12 kHz streams, +1.5 s shift, +-10 s window'),
    # BASELINE.json configs[1]: the single-GPU configuration the metric is quoted on
    'config2': dict(events=2000, duration=1800.0, window=60.0, min_len=1.0, max_len=4.0, shift=1.5, scaling='weak',
                    text='2000 events, 2x30 min 12 kHz streams, +-60 s window'),
    # BASELINE.json configs[2]: the 8-GPU target configuration (events sharded: strong scaling)
    'config3': dict(events=10000, duration=5400.0, window=120.0, min_len=1.0, max_len=4.0, shift=1.5, scaling='strong',
                    text='10000 events, 2x90 min 12 kHz streams, +-120 s window, event-sharded'),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------
class HostStream(object):
    """A normalised stream on the host (what WavStream.data holds) without any GPU state."""

    def __init__(self, pcm, sample_type):
        s = object.__new__(WavStream)
        s._handle = None

        class _Mem(object):
            pass
        from sushi_b200.wavstream import decode_downmix
        raw = np.ascontiguousarray(pcm, '<i2').view(np.uint8)
        mem = _Mem()
        mem.framerate, mem.channels_count, mem.sample_width, mem.frame_size = SAMPLE_RATE, 1, 2, 2
        mem.frames_count = raw.size // 2
        pos = [0]

        def readframes(count):
            a = pos[0]
            b = min(a + count * 2, raw.size)
            pos[0] = b
            return decode_downmix(raw[a:b].tobytes(), 2, 1)
        mem.readframes = readframes
        s._load(mem, SAMPLE_RATE, sample_type)
        self.data, self.sample_count, self.padding_size = s.data, s.sample_count, s.padding_size
        self.sample_type = sample_type


def make_inputs(wl, sample_type, n_lists, seed=2):
    t0 = time.time()
    src_pcm, dst_pcm = synth.make_pair(wl['duration'], seed, wl['shift'])
    src = HostStream(src_pcm, sample_type)
    dst = HostStream(dst_pcm, sample_type)
    lists = [synth.make_events(wl['events'], wl['duration'], seed + 1000 * r, wl['min_len'], wl['max_len'])
             for r in range(n_lists)]
    log('[bench] inputs ready in %.1f s' % (time.time() - t0))
    return src, dst, lists


def algorithmic_bytes(tlen, nlags, bytes_per_sample):
    """one template read + one pass over the search span + (diff, idx) out."""
    tlen = np.asarray(tlen, np.float64)
    nlags = np.asarray(nlags, np.float64)
    return float(np.sum(bytes_per_sample * tlen + bytes_per_sample * (nlags + tlen - 1) + 16))


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init():
    import cv2
    cv2.setNumThreads(1)       # matchTemplate does not scale with threads on 1-row images


def _cpu_run(chunk):
    src, dst = _CPU['src'], _CPU['dst']
    out = []
    for (a, b, c, w) in chunk:
        d, t = dst.find_substream(src.get_substream(a, b), c, w)
        out.append((float(d), t))
    return out


def effective_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                if txt[0] != 'max':
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                quota = float(txt[0])
                period = float(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
                if quota > 0:
                    n = min(n, max(1, int(quota / period)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def cpu_events_per_s(src, dst, starts, ends, window, budget_s, cores=None):
    """Throughput of the reference's CPU path on a bounded sample of the workload's events."""
    import multiprocessing as mp
    from oracle.ref_matcher import RefStream
    cores = cores or effective_cores()
    _CPU['src'] = RefStream(src.data, SAMPLE_RATE, src.padding_size, src.sample_count)
    _CPU['dst'] = RefStream(dst.data, SAMPLE_RATE, dst.padding_size, dst.sample_count)
    _cpu_init()
    order = np.linspace(0, len(starts) - 1, min(len(starts), 8)).astype(int)
    t0 = time.perf_counter()
    _cpu_run([(starts[order[0]], ends[order[0]], starts[order[0]], window)])       # warm-up
    t1 = time.perf_counter()
    _cpu_run([(starts[i], ends[i], starts[i], window) for i in order[1:3]])
    per_event = max((time.perf_counter() - t1) / 2, 1e-4)
    m = int(max(cores, min(len(starts), cores * budget_s / per_event)))
    m = min(m, len(starts))
    sel = np.linspace(0, len(starts) - 1, m).astype(int)
    items = [(starts[i], ends[i], starts[i], window) for i in sel]
    nchunk = min(len(items), cores * 4)
    chunks = [items[i::nchunk] for i in range(nchunk)]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_run, [[items[0]]] * cores)                                   # spin the workers up
        t0 = time.perf_counter()
        res = pool.map(_cpu_run, chunks)
        wall = time.perf_counter() - t0
    first = res[0][0]
    return m / wall, cores, m, per_event, first


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix='clocks_', suffix='.csv')
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '--query-gpu=' + self.FIELDS, '--format=csv,noheader,nounits', '-lms', '50',
                 '-i', str(gpu_index)], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            p = [x.strip() for x in line.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v == 'Active':
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def pinned_copy(arr):
    from sushi_b200 import _native
    out = _native.pinned_empty(arr.shape, arr.dtype)
    np.copyto(out, arr)
    return out


def run_b200(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('bench.py --gpus %d must be launched with torch.distributed.run (WORLD_SIZE=%d)' % (args.gpus, world))
    wl = WORKLOADS[args.workload]
    stype = args.sample_type
    bps = 1 if stype == 'uint8' else 4
    strong = wl['scaling'] == 'strong'

    src_h, dst_h, lists = make_inputs(wl, stype, 1 if strong else world)
    if strong:
        starts_all, ends_all = lists[0]
        my = slice(*parallel.shard_bounds(len(starts_all), world, rank))      # contiguous shard
        starts, ends = starts_all[my], ends_all[my]
        total_events = len(starts_all)
    else:
        starts, ends = lists[rank]
        total_events = wl['events'] * world
    centers = starts.copy()
    windows = np.full(len(starts), wl['window'])

    # ---- CPU baseline first (fork before CUDA is initialised in this process) -------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, m, per_event, _ = cpu_events_per_s(src_h, dst_h, lists[0][0], lists[0][1], wl['window'], args.cpu_budget)
        cpu = {'value': round(v, 3), 'unit': 'events/s', 'cores': cores, 'kind': 'port',
               'sample': '%d of the %d events of %s, evenly spaced, one process per core, cv2 threads=1; '
                         '1-core calibration %.1f ms/event' % (m, wl['events'], args.workload, per_event * 1e3)}
        log('[bench] cpu baseline: %.1f events/s on %d cores (%d events)' % (v, cores, m))

    from sushi_b200 import _native
    lib = _native.lib(local_rank)
    if args.block:
        _native.check(lib.sb_set_block_size(args.block))
    if args.chunk:
        _native.check(lib.sb_set_chunk_items(args.chunk))
    if args.engine >= 0:
        _native.check(lib.sb_set_engine(args.engine))
    if args.hop_mode >= 0:
        _native.check(lib.sb_set_hop_mode(args.hop_mode))
    if args.premac_mode >= 0:
        _native.check(lib.sb_set_premac_mode(args.premac_mode))
    if args.epilogue > 0:
        _native.check(lib.sb_set_epilogue(args.epilogue))

    dist = torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        ext = torch.cuda.ExternalStream(lib.sb_get_stream(), device=torch.device('cuda', local_rank))

    # page-locked host copies (the e2e leg copies from these every step)
    src_p, dst_p = pinned_copy(src_h.data), pinned_copy(dst_h.data)
    n_src, n_dst = src_p.shape[1], dst_p.shape[1]

    # resident raw streams + the integer plan (value leg)
    src0 = WavStream.from_array(src_p, SAMPLE_RATE, src_h.padding_size, src_h.sample_count)
    dst0 = WavStream.from_array(dst_p, SAMPLE_RATE, dst_h.padding_size, dst_h.sample_count)
    toff, tlen, lag0, nlags, t0s = dst0.plan_queries(src0, starts, ends, centers, windows)
    count = len(toff)
    step_bytes = algorithmic_bytes(tlen, nlags, bps)

    if world > 1:
        src_t = torch.empty(n_src, dtype=torch.uint8 if stype == 'uint8' else torch.float32, device='cuda')
        dst_t = torch.empty_like(src_t) if n_dst == n_src else torch.empty(n_dst, dtype=src_t.dtype, device='cuda')
        if rank == 0:
            src_t.copy_(torch.from_numpy(src_p[0]))
            dst_t.copy_(torch.from_numpy(dst_p[0]))
        maxc = max(parallel.shard_sizes(total_events, world)) if strong else count
        pad_diff = torch.zeros(maxc, dtype=torch.float32, device='cuda')
        pad_idx = torch.zeros(maxc, dtype=torch.int64, device='cuda')
        all_diff = torch.empty(world * maxc, dtype=torch.float32, device='cuda')
        all_idx = torch.empty(world * maxc, dtype=torch.int64, device='cuda')
        torch.cuda.synchronize()
        src_ptr, dst_ptr = src_t.data_ptr(), dst_t.data_ptr()
        d_diff_ptr, d_idx_ptr = pad_diff.data_ptr(), pad_idx.data_ptr()
    else:
        src_ptr, dst_ptr = src0.device_ptr, dst0.device_ptr
        import ctypes
        pd, pi = ctypes.c_void_p(), ctypes.c_void_p()
        _native.check(lib.sb_device_alloc(4 * count, ctypes.byref(pd)))
        _native.check(lib.sb_device_alloc(8 * count, ctypes.byref(pi)))
        d_diff_ptr, d_idx_ptr = pd.value, pi.value

    def step_device():
        """value leg: raw streams resident in HBM -> per-event (diff, idx) in HBM."""
        if world > 1:
            with torch.cuda.stream(ext):
                dist.broadcast(src_t, 0)          # the one NCCL broadcast of the streams
                dist.broadcast(dst_t, 0)
        s = WavStream.from_device(src_ptr, n_src, stype, SAMPLE_RATE, src_h.padding_size, src_h.sample_count)
        d = WavStream.from_device(dst_ptr, n_dst, stype, SAMPLE_RATE, dst_h.padding_size, dst_h.sample_count)
        d.find_planned_device(s, toff, tlen, lag0, nlags, d_diff_ptr, d_idx_ptr)
        if world > 1:
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(all_diff, pad_diff)      # per-event results to every rank
                dist.all_gather_into_tensor(all_idx, pad_idx)
        s.close()
        d.close()

    def step_e2e():
        """e2e leg: the public API from host buffers, results back on the host."""
        s = WavStream.from_array(src_p, SAMPLE_RATE, src_h.padding_size, src_h.sample_count)
        d = WavStream.from_array(dst_p, SAMPLE_RATE, dst_h.padding_size, dst_h.sample_count)
        diffs, times = d.find_substream_batch(s, starts, ends, centers, windows)
        s.close()
        d.close()
        return diffs, times

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        _native.check(lib.sb_sync())

    def timed_device(k):
        import ctypes
        barrier()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(k):
                step_device()
            e1.record(ext)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())
        _native.check(lib.sb_timer_start())
        for _ in range(k):
            step_device()
        ms = ctypes.c_float()
        _native.check(lib.sb_timer_stop(ctypes.byref(ms)))
        return float(ms.value)

    # ---- correctness of what is being timed: known shift recovered --------------------
    diffs, times = step_e2e()
    shifts = times - starts
    ok = (ends + wl['shift'] < wl['duration']) & (starts + wl['shift'] > 0)
    bad = int(np.sum(np.abs(shifts[ok] - wl['shift']) > 1.0 / SAMPLE_RATE + 1e-9))
    if bad:
        log('[bench] WARNING: %d of %d events did not recover the known shift' % (bad, int(ok.sum())))

    # ---- warm-up, then the timed regions -----------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    lib.sb_profile_reset()
    lib.sb_profile_enable(1)
    ms_total = timed_device(args.steps)
    lib.sb_profile_enable(0)
    clocks = sampler.stop() if sampler else None
    launches = int(lib.sb_launch_count())
    prof = {}
    import ctypes
    for name in lib.sb_profile_names().decode().split(','):
        if not name:
            continue
        ms, n = ctypes.c_double(), ctypes.c_int64()
        lib.sb_profile_get(name.encode(), ctypes.byref(ms), ctypes.byref(n))
        if n.value or ms.value:
            prof[name] = (ms.value, n.value)

    # e2e (host buffers, public API), wall clock bracketed by device syncs, max over ranks
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        from sushi_b200 import _native as nat
        ms_step = ms_total / args.steps
        value = total_events / (ms_step / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback 6.65 TB/s'
        # dominant kernel class by accumulated device time inside the timed steps
        dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
        dom_name, (dom_ms, dom_n) = dom
        dom_ms_step = dom_ms / args.steps
        achieved = step_bytes / (dom_ms_step / 1e3) / 1e9 if dom_ms_step > 0 else 0.0
        # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this same
        # workload (profiles/traffic.json, written by tools/ncu_traffic.py); null when no capture matches
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            key = '%s/%s/%s/B%d' % (args.workload, stype, dom_name, lib.sb_get_block_size())
            if key in tr:
                traffic = tr[key]['dram_bytes_per_launch']
        except (OSError, ValueError, KeyError):
            pass
        roof = {'bound': 'hbm', 'kernel': dom_name, 'achieved': round(achieved, 2), 'peak': peak, 'unit': 'GB/s',
                'frac': round(achieved / peak, 5), 'traffic': traffic, 'peak_source': peak_src,
                'launches_per_step': dom_n / args.steps, 'avg_launch_ms': round(dom_ms / max(dom_n, 1), 5),
                'algorithmic_bytes_per_step': step_bytes,
                'whole_step': {'achieved': round(step_bytes / (ms_step / 1e3) / 1e9, 2),
                               'frac': round(step_bytes / (ms_step / 1e3) / 1e9 / peak, 5)},
                'kernel_ms_per_step': {k: round(v[0] / args.steps, 4) for k, v in sorted(prof.items())}}
        h2d = int(src_p.nbytes + dst_p.nbytes + count * 56)
        d2h = int(count * 12)
        e2e_value = total_events * args.steps / e2e_s
        line = {
            'metric': 'subtitle_events_per_s', 'value': round(value, 2), 'unit': 'events/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': round(ms_step, 4), 'ms_per_event': round(ms_step * world / total_events, 6),
            'higher_is_better': True, 'scaling': wl['scaling'], 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload + ': ' + wl['text'], 'events_per_gpu': count, 'sample_type': stype,
                       'sample_rate': SAMPLE_RATE, 'window_s': wl['window'], 'lag_block': lib.sb_get_block_size(), 'epilogue': lib.sb_get_epilogue(), 'spectra': 'f32', 'engine': {0: 'cufft', 1: 'fused', 2: 'fused_packed', 4: 'fused_packed_pair', 5: 'fused_packed_single'}[lib.sb_get_engine()],
                       'parallelism': 'events x%d' % world,
                       'l2': 'working set > L2: block spectra %.0f MB + running sums %.0f MB per stream, rebuilt every step'
                             % (n_dst * 8 / 1e6, n_dst * 16 / 1e6),
                       'step': 'running sums + block spectra + all queries' + (' + NCCL broadcast/all-gather' if world > 1 else '')},
            'roofline': roof,
            'e2e': {'value': round(e2e_value, 2), 'unit': 'events/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': round(e2e_s / args.steps * 1e3, 3)},
            'gpu_launches': launches,
            'clocks': clocks,
            'shift_check': {'events_checked': int(ok.sum()), 'mismatches': bad},
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    src_h, dst_h, lists = make_inputs(wl, args.sample_type, 1)
    starts, ends = lists[0]
    vals = []
    budget = max(2.0, min(args.cpu_budget, 60.0 / max(args.steps + args.warmup, 1)))
    cores = m = 0
    for i in range(args.warmup + args.steps):
        v, cores, m, per_event, _ = cpu_events_per_s(src_h, dst_h, starts, ends, wl['window'], budget)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'subtitle_events_per_s', 'value': round(value, 3), 'unit': 'events/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(m / value * 1e3, 3), 'higher_is_better': True, 'scaling': wl['scaling'],
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload + ': ' + wl['text'], 'sample_type': args.sample_type,
                   'sample_rate': SAMPLE_RATE, 'window_s': wl['window']},
        'cpu_baseline': {'value': round(value, 3), 'unit': 'events/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d of the %d events per step, evenly spaced, one process per core, cv2 threads=1'
                                   % (m, wl['events'])},
        'e2e': {'value': round(value, 3), 'unit': 'events/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS))
    ap.add_argument('--sample-type', default='uint8', choices=['uint8', 'float32'])
    ap.add_argument('--cpu-budget', type=float, default=12.0, help='seconds of wall clock for the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--block', type=int, default=0, help='lag-block size override')
    ap.add_argument('--chunk', type=int, default=0, help='items per launch override')
    ap.add_argument('--premac-mode', type=int, default=-1, help='blocked multiply kernel: 0 by template length (default), 1 never, 2 always')
    ap.add_argument('--hop-mode', type=int, default=-1, help='fused engine geometry: 1 hop B (default), 2 hop B/2, 0 cost rule per batch')
    ap.add_argument('--epilogue', type=int, default=0, help='body variant of the packed kernels on uint8 streams: 2 trimmed (default), 1 first version')
    ap.add_argument('--engine', type=int, default=-1, help='0: cuFFT pipeline, 1: fused kernel, 2: packed fused kernels (default), 4 / 5: always / never pairs of lag blocks')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
