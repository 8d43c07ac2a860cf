"""The LOGIC of the packed match kernels (sushi_b200/csrc/sb_fused2.cu) on the CPU.

tests/emu/ compiles the kernels' own source for the host (g++, -DSB_EMULATE: host stand-ins for the built-in
variables, the intrinsics and the inline-PTX wrappers of sb_ptx.cuh) and runs CTAs with one OS thread per warp
whose 32 lanes are fibers (one OS thread per CUDA thread under ThreadSanitizer, tests/emu/run_tsan.py).  This is test infrastructure like oracle/: it is never loaded by the product path, and it proves nothing
about races, fences, alignment rules of the copy engine or speed -- tests/test_gpu_*.py do that on a B200.  What it
does pin, without a GPU: the index algebra (quad rows -> packing -> FFT passes -> epilogue), the bookkeeping of the
pair kernel (mbarrier phases, which thread parks what in which tensor-memory columns, the last group of
a query holding fewer lag blocks, spectrum rows past the end of the stream), the candidate logic of both screening
loops, and the first-index argmin -- against the fp64 closed form of TM_SQDIFF_NORMED and across kernels bit for
bit.  Spectrum rows and running sums are prepared here in NumPy the way k_forward_quad / the scan kernels define
them (layout: tests/test_packed_layout_model.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.ref_matcher import sqdiff_normed_fp64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, 'tests', 'emu')
B = 16384
Q4 = B // 4


class QueryDesc(ctypes.Structure):            # sb_internal.h
    _fields_ = [(n, ctypes.c_int64) for n in ('toff', 'tlen', 'lag0', 'nlags', 'itemBase', 'partBase', 'curveOff', 'groupBase')] + \
               [(n, ctypes.c_int32) for n in ('P', 'k0', 'nk', 'orig')]


@pytest.fixture(scope='module')
def emu():
    src = [os.path.join(EMU, f) for f in ('emu_driver.cpp', 'emu_cuda.h', 'emu_ptx.h')] + \
          [os.path.join(ROOT, 'sushi_b200', 'csrc', f) for f in ('sb_fused2.cu', 'sb_fused_common.cuh', 'sb_fft_smem.cuh', 'sb_internal.h')]
    out = os.path.join(EMU, '_build', 'libsb_emu.so')
    if os.environ.get('SB_EMU_LIB'):                      # a prebuilt library (scratch builds of a kernel under development)
        lib = ctypes.CDLL(os.environ['SB_EMU_LIB'])
        assert lib.emu_query_desc_bytes() == ctypes.sizeof(QueryDesc)
        return lib
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
        subprocess.check_call(['g++', '-std=c++20', '-O1', '-pthread', '-DSB_EMULATE', '-I', EMU,
                               '-I', os.path.join(ROOT, 'sushi_b200', 'csrc'), '-I', os.path.join(ROOT, 'include'), '-I', cuda_inc,
                               '-shared', '-fPIC', src[0], '-o', out])
    lib = ctypes.CDLL(out)
    assert lib.emu_query_desc_bytes() == ctypes.sizeof(QueryDesc)
    return lib


def aligned(n, dtype, fill=0):
    """1-D array of n elements whose data pointer is a multiple of 128 (the copy engine wants 16)."""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + 256, np.uint8)
    off = (-raw.ctypes.data) % 128
    a = raw[off:off + n * item].view(dtype)
    a[:] = fill
    return a


def qa(i):
    return (i >> 8) * 512 + (i & 255) if i < Q4 else 2 * Q4


def quad_rows(blocks, row_floats):
    """blocks: (rows, 2B) real -> rows in the quad layout of sb_fused2.cu."""
    X = np.fft.rfft(blocks.astype(np.float64), axis=1)          # bins 0 .. B
    out = aligned(blocks.shape[0] * row_floats, np.float32).reshape(blocks.shape[0], row_floats // 4, 4)
    i = np.arange(Q4 + 1)
    pa = np.array([qa(int(v)) for v in i])
    pm = np.where(i < Q4, pa + 256, 2 * Q4 + 1)
    out[:, pa, 0], out[:, pa, 1], out[:, pa, 2], out[:, pa, 3] = X[:, i].real, X[:, i + B // 2].real, X[:, i].imag, X[:, i + B // 2].imag
    out[:, pm, 0], out[:, pm, 1], out[:, pm, 2], out[:, pm, 3] = X[:, B - i].real, X[:, B // 2 - i].real, X[:, B - i].imag, X[:, B // 2 - i].imag
    return out.reshape(blocks.shape[0], row_floats)


def prefix_sums(x):
    p = aligned(2 * (x.size + 1), np.float64).reshape(x.size + 1, 2)
    xf = x.astype(np.float64)
    p[1:, 0], p[1:, 1] = np.cumsum(xf), np.cumsum(xf * xf)
    return p


class Case(object):
    """One image stream, one template stream, a list of queries (toff, n, lag0, nlags)."""

    def __init__(self, lib, img, src, queries, dtype):
        self.lib, self.queries, self.dtype = lib, queries, dtype
        rf = lib.emu_quad_row_floats()
        n_img = img.size
        self.img = aligned(n_img + 64, dtype)
        self.img[:n_img] = img
        self.n_img = n_img
        self.ipfx, self.tpfx = prefix_sums(img), prefix_sums(src)
        centre = (lambda s, c: np.rint(s / c)) if dtype == np.uint8 else (lambda s, c: np.float64(np.float32(s / c)))
        a = centre(self.ipfx[n_img, 0], n_img)
        self.nblk = (n_img + B - 1) // B
        blocks = np.zeros((self.nblk, 2 * B))
        for k in range(self.nblk):
            seg = img[k * B:k * B + 2 * B].astype(np.float64)
            blocks[k, :seg.size] = seg - a
        self.Xhat = quad_rows(blocks, rf)
        parts = []
        for (toff, n, lag0, nlags) in queries:
            t = src[toff:toff + n].astype(np.float64)
            b = centre(t.sum(), n)
            for p in range((n + B - 1) // B):
                row = np.zeros(2 * B)
                seg = t[p * B:(p + 1) * B]
                row[:seg.size] = seg - b
                parts.append(row)
        self.That = quad_rows(np.array(parts), rf)
        self.src = src

    def run(self, kernel, epi, curves, records=None):
        """records: the third body hands its selected runs to k_finish_runs (default whenever epi == 3 and no curve
        is asked for, like the library); False keeps everything in the match kernel."""
        records = (epi == 3 and not curves) if records is None else records
        group = {0: 1, 1: 2}[kernel]
        desc = (QueryDesc * len(self.queries))()
        items = parts = groups = curve = 0
        cta_query = []
        for q, (toff, n, lag0, nlags) in enumerate(self.queries):
            d = desc[q]
            d.toff, d.tlen, d.lag0, d.nlags = toff, n, lag0, nlags
            d.P, d.k0 = (n + B - 1) // B, lag0 // B
            d.nk = (lag0 + nlags - 1) // B - d.k0 + 1
            d.itemBase, d.partBase, d.curveOff, d.groupBase, d.orig = items, parts, curve, groups, q
            ng = (d.nk + group - 1) // group
            cta_query += [q] * ng
            items += d.nk
            parts += d.P
            groups += ng
            curve += nlags
        cta_query = np.array(cta_query, np.int32)
        keys = np.full(len(self.queries), 0xffffffffffffffff, np.uint64)
        cur = np.full(curve, np.nan, np.float32) if curves else None
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        recs = counts = None
        if records:
            recs = np.full(len(cta_query) * self.lib.emu_run_slots() * self.lib.emu_run_record_bytes(), 0xCD, np.uint8)
            counts = np.full(len(cta_query), -1, np.int32)
        rc = self.lib.emu_run(kernel, epi, int(self.dtype == np.uint8), vp(self.That), ctypes.c_int64(0), vp(self.Xhat), ctypes.c_int64(self.nblk),
                              vp(self.img), ctypes.c_int64(self.n_img), vp(self.ipfx), vp(self.tpfx), ctypes.byref(desc), vp(cta_query),
                              ctypes.c_int64(0), len(cta_query), vp(keys), vp(cur) if curves else None,
                              vp(recs) if records else None, vp(counts) if records else None)
        assert rc == 0, 'emulation reported %d errors (see stderr)' % rc
        if records:
            assert (counts >= 0).all() and (counts <= self.lib.emu_run_slots()).all()
            self.last_record_counts = counts.copy()
            self.lib.emu_finish_runs(vp(recs), vp(counts), len(cta_query), ctypes.byref(desc), vp(self.ipfx), ctypes.c_int64(self.n_img),
                                     vp(self.tpfx), vp(keys))
        diff = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
        idx = (keys & np.uint64(0xffffffff)).astype(np.int64)
        return diff, idx, cur

    def truth(self):
        out = []
        for (toff, n, lag0, nlags) in self.queries:
            out.append(sqdiff_normed_fp64(self.img[lag0:lag0 + nlags + n - 1], self.src[toff:toff + n]))
        return out


def programme(n, seed):
    rng = np.random.default_rng(seed)
    x = np.convolve(rng.standard_normal(n + 8), np.hanning(9), 'valid') * np.repeat(rng.uniform(0.2, 1.0, n // 2400 + 1), 2400)[:n]
    return np.clip(np.rint(128 + 70 * x), 0, 255).astype(np.uint8)


@pytest.fixture(scope='module')
def case_u8(emu):
    n_img = 6 * B - 5000                                     # 6 block rows, the last one short
    img = programme(n_img, 1)
    rng = np.random.default_rng(2)
    src = np.clip(np.roll(img, -700).astype(np.int32) + rng.integers(-5, 6, n_img), 0, 255).astype(np.uint8)   # src(t) = img(t + 700) + noise
    queries = [(30000, 20000, B + 300, 4 * B - 17000),       # P = 2, 4 lag blocks, starts and ends inside a block; match at lag 30700
               (1000, 5000, 100, 20000),                      # P = 1, 2 lag blocks; match at lag 1700
               (8000, 40000, n_img - 40000 - 30000, 30001)]   # P = 3, 3 lag blocks, runs to the very end of the stream
    return Case(emu, img, src, queries, np.uint8)


def test_emulated_kernels_match_the_closed_form_and_each_other(case_u8):
    c = case_u8
    truth = c.truth()
    ref = None
    for kernel in (0, 1):                                     # one CTA per lag block / pair
        for epi in (1, 3):
            d_c, i_c, cur = c.run(kernel, epi, curves=True)   # every lag through the exact path
            d_s, i_s, _ = c.run(kernel, epi, curves=False)    # screening decides which lags are evaluated
            off = 0
            for q, t in enumerate(truth):
                got = cur[off:off + t.size]
                off += t.size
                assert not np.isnan(got).any()
                assert np.abs(got - t).max() <= 3e-6
                assert i_c[q] == int(got.argmin()) and d_c[q] == got.min()
                assert abs(int(got.argmin()) - int(t.argmin())) <= 1
            assert np.array_equal(d_s, d_c) and np.array_equal(i_s, i_c)     # screening lost nothing
            if ref is None:
                ref = (d_c, i_c, cur)
            else:                                                            # same arithmetic in the same order
                assert np.array_equal(ref[0], d_c) and np.array_equal(ref[1], i_c) and np.array_equal(ref[2], cur)
    assert ref[1][0] == 30700 - (B + 300) and ref[1][1] == 1700 - 100


def test_emulated_degenerate_blocks(emu):
    """Silence inside programme material, a zero template, an exact copy: saturated blocks must evaluate every lag
    under both screening loops and return the FIRST index of equal minima."""
    n_img = 3 * B
    img = programme(n_img, 5)
    img[9000:31000] = 0
    src = img.copy()
    src[40000:41000] = 0
    queries = [(12000, 6000, 0, 34001),        # template of silence (zero template): every value 1 -> index 0
               (100, 5000, 8000, 20000),        # lags 9000..26000 see silent windows (value 1), the rest programme
               (33000, 5000, 0, 2 * B + 1000),  # exact copy at lag 33000: value 0
               (40000, 1000, 100, 1000)]        # zero template against programme
    c = Case(emu, img, src, queries, np.uint8)
    truth = c.truth()
    ref = None
    for kernel in (0, 1):
        for epi in (1, 3):
            d, i, _ = c.run(kernel, epi, curves=False)
            for q, t in enumerate(truth):
                assert i[q] == int(t.argmin()), (kernel, epi, q)
                assert abs(float(d[q]) - float(t.min())) <= 3e-6
            ref = ref or (d, i)
            assert np.array_equal(ref[0], d) and np.array_equal(ref[1], i)
    assert ref[0][0] == 1.0 and ref[1][0] == 0 and ref[0][2] <= 1e-6 and ref[1][2] == 33000 and ref[0][3] == 1.0 and ref[1][3] == 0


def test_emulated_float32_stream(emu):
    n_img = 4 * B - 3000
    rng = np.random.default_rng(9)
    img = (programme(n_img, 3).astype(np.float32) / 255.0).astype(np.float32)
    src = (np.roll(img, -300) + rng.normal(0, 0.01, n_img)).astype(np.float32)
    c = Case(emu, img, src, [(20000, 18000, 5, 2 * B + 5000)], np.float32)
    t = c.truth()[0]
    ref = None
    for kernel in (0, 1):
        d, i, cur = c.run(kernel, 1, curves=True)
        assert np.abs(cur - t).max() <= 3e-6 and i[0] == int(cur.argmin()) == 20300 - 5
        ref = ref or (d, i, cur)
        assert np.array_equal(ref[2], cur) and ref[0][0] == d[0]


def test_emulated_forward_kernel_writes_the_quad_rows(emu, case_u8):
    """k_forward_quad (block spectra of a stream) in emulation against the NumPy rows the other tests feed the match
    kernels: agreement to fp32 FFT rounding, and matching on the kernel's own rows gives the same answer."""
    c = case_u8
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rf = emu.emu_quad_row_floats()
    out = aligned(c.nblk * rf, np.float32)
    assert emu.emu_forward_blocks(1, vp(c.img), ctypes.c_int64(c.n_img), vp(c.ipfx), ctypes.c_int64(0), c.nblk, vp(out)) == 0
    rows = out.reshape(c.nblk, rf)
    used = 2 * Q4 + 2
    want0 = c.Xhat[:, :used * 4]
    scale = np.abs(want0).max()
    assert np.abs(rows[:, :used * 4] - want0).max() <= 2e-6 * scale
    d_ref, i_ref, _ = c.run(1, 3, curves=False)
    keep = c.Xhat
    try:
        c.Xhat = rows
        d, i, _ = c.run(1, 3, curves=False)
    finally:
        c.Xhat = keep
    assert np.array_equal(i, i_ref) and np.abs(d - d_ref).max() <= 1e-6


def test_emulated_edge_geometry(emu):
    """Templates of 3, B and B + 1 samples, ranges of a single lag, ranges whose first / last lag is a block's first
    / last lag, the last lags of the stream (staged windows clamped at the end of the allocation): every kernel,
    both screening loops."""
    n_img = 3 * B + 777
    img = programme(n_img, 21)
    rng = np.random.default_rng(22)
    src = np.clip(np.roll(img, -50).astype(np.int32) + rng.integers(-3, 4, n_img), 0, 255).astype(np.uint8)
    queries = [(500, 3, 0, 4000),                              # three samples
               (1000, B, 900, 300),                            # exactly one partition; match at lag 1050
               (1000, B + 1, 2 * B - 10, 20),                  # two partitions, the second holds one sample; straddles a block edge
               (7000, 9000, B, 1),                             # one lag, the first of a block
               (7000, 9000, 2 * B - 1, 1),                     # one lag, the last of a block
               (20000, 12000, n_img - 12000 - 5000, 5001),     # up to the last possible lag of the stream
               (123, 300, 3 * B - 3, 481)]                     # into the short last block, up to the last lag
    c = Case(emu, img, src, queries, np.uint8)
    truth = c.truth()
    ref = None
    for kernel in (0, 1):
        for epi in (1, 3):
            d, i, cur = c.run(kernel, epi, curves=True)
            d_s, i_s, _ = c.run(kernel, epi, curves=False)
            assert np.array_equal(d, d_s) and np.array_equal(i, i_s)
            off = 0
            for q, t in enumerate(truth):
                got = cur[off:off + t.size]
                off += t.size
                assert np.abs(got - t).max() <= 3e-6, (kernel, epi, q)
                assert i[q] == int(got.argmin()) and d[q] == got.min()
            ref = ref or (d, i, cur)
            assert np.array_equal(ref[0], d) and np.array_equal(ref[1], i) and np.array_equal(ref[2], cur)
    assert ref[1][1] == 1050 - 900


def test_emulated_variants_reproduce_the_reference_golden(emu, golden_matcher):
    """The reference's own outputs (tests/golden/matcher.npz: find_substream of /root/reference/wav.py over cv2 on
    12 queries) through the emulated default kernel (pairs of lag blocks, body 3 with and without records) within north_star's
    tolerances: shift +-1 sample, diff 1e-5."""
    from tests.helpers import oracle_stream_from_pcm
    g = golden_matcher
    rs = oracle_stream_from_pcm(g['src_pcm'], 12000, 1, 12000, 'uint8')
    rd = oracle_stream_from_pcm(g['dst_pcm'], 12000, 1, 12000, 'uint8')
    clip = lambda v, lo, hi: max(min(v, hi), lo)
    keep = list(range(len(g['queries'])))
    queries, t0s = [], []
    for (a, b, c, w) in g['queries'][keep]:
        toff = rs.sample_for_time(a)
        n = rs.sample_for_time(b) - toff
        start = clip(c - w, -10, rd.duration_seconds)                 # wav.py:178-182
        end = clip(c + w, 0, rd.duration_seconds + 10)
        lag0 = rd.sample_for_time(start)
        nlags = rd.sample_for_time(end) + n - lag0 - n + 1
        queries.append((toff, n, lag0, nlags))
        t0s.append(start)
    case = Case(emu, rd.data[0], rs.data[0], queries, np.uint8)
    for epi, records in ((3, True), (3, False)):
        d, i, _ = case.run(1, epi, curves=False, records=records)
        times = np.array(t0s) + i / 12000.0
        assert np.abs(d - g['diff_uint8'][keep]).max() <= 1e-5, np.abs(d - g['diff_uint8'][keep]).max()
        assert np.abs(times - g['time_uint8'][keep]).max() <= 1.0 / 12000 + 1e-9


def test_emulated_config1_against_the_live_oracle(emu):
    """BASELINE config 1 (100 events, 2 x 60 s streams with a constant +1.5 s shift, +-10 s window) through the
    emulated kernels -- pairs of lag blocks with the first body and with the third (default) one -- against the
    oracle's find_substream (cv2) on every event."""
    from sushi_b200 import synth
    from tests.helpers import oracle_stream_from_pcm
    src_pcm, dst_pcm = synth.make_pair(60.0, 2, 1.5)
    rs = oracle_stream_from_pcm(src_pcm, 12000, 1, 12000, 'uint8')
    rd = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, 'uint8')
    starts, ends = synth.make_events(100, 60.0, 1002, 1.0, 4.0)
    clip = lambda v, lo, hi: max(min(v, hi), lo)
    queries, t0s, want = [], [], []
    for a, b in zip(starts, ends):
        toff = rs.sample_for_time(a)
        n = rs.sample_for_time(b) - toff
        start = clip(a - 10.0, -10, rd.duration_seconds)
        end = clip(a + 10.0, 0, rd.duration_seconds + 10)
        lag0 = rd.sample_for_time(start)
        queries.append((toff, n, lag0, rd.sample_for_time(end) - lag0 + 1))
        t0s.append(start)
        want.append(rd.find_substream(rs.get_substream(a, b), a, 10.0))
    want_d = np.array([w[0] for w in want], np.float64)
    want_t = np.array([w[1] for w in want])
    case = Case(emu, rd.data[0], rs.data[0], queries, np.uint8)
    for kernel, epi in ((1, 1), (1, 3)):
        d, i, _ = case.run(kernel, epi, curves=False)
        times = np.array(t0s) + i / 12000.0
        assert np.abs(d - want_d).max() <= 1e-5, (kernel, epi, np.abs(d - want_d).max())
        assert np.abs(times - want_t).max() <= 1.0 / 12000 + 1e-9
        ok = ends + 1.5 < 60.0
        assert np.abs((times - starts)[ok] - 1.5).max() <= 1.0 / 12000 + 1e-9        # the known answer


@pytest.mark.parametrize('seed', [101, 102, 103, 104])
def test_emulated_random_queries_all_variants_agree(emu, seed):
    """Random template lengths (all residues of the window alignment), random ranges: one CTA per lag block with the
    first body against both kernels with the third body, bit for bit, and against the closed form."""
    rng = np.random.default_rng(seed)
    n_img = int(rng.integers(2 * B + 100, 5 * B))
    img = programme(n_img, seed)
    src = np.clip(np.roll(img, -int(rng.integers(0, 4000))).astype(np.int32) + rng.integers(-4, 5, n_img), 0, 255).astype(np.uint8)
    queries = []
    for _ in range(5):
        n = int(rng.integers(1, min(3 * B, n_img - 10)))
        toff = int(rng.integers(0, n_img - n + 1))
        lag0 = int(rng.integers(0, n_img - n + 1))
        nlags = int(rng.integers(1, n_img - n - lag0 + 2))
        queries.append((toff, n, lag0, min(nlags, 2 * B + 5000)))
    c = Case(emu, img, src, queries, np.uint8)
    truth = c.truth()
    d0, i0, _ = c.run(0, 1, curves=False)
    for kernel, epi in ((1, 3), (0, 3), (1, 1)):
        d, i, _ = c.run(kernel, epi, curves=False)
        assert np.array_equal(d, d0) and np.array_equal(i, i0), (seed, kernel, epi, queries)
    for q, t in enumerate(truth):
        assert abs(float(d0[q]) - float(t.min())) <= 3e-6 and abs(int(i0[q]) - int(t.argmin())) <= 1, (seed, q, queries[q])


def test_emulated_many_partitions_do_not_overrun_the_special_area(emu):
    """A template of 13 partitions through the packed kernels (what sb_set_premac_mode(1) allows): 2P + G - 1 rows
    do not fit the 32-row staging area of the self-mirrored quad, so these CTAs must read it from L2 -- same
    answers as the first kernel body (which always does), bit for bit, and the closed form's."""
    n_img = 16 * B
    img = programme(n_img, 31)
    rng = np.random.default_rng(32)
    src = np.clip(np.roll(img, -500).astype(np.int32) + rng.integers(-4, 5, n_img), 0, 255).astype(np.uint8)
    n = 12 * B + 1234                                           # P = 13
    c = Case(emu, img, src, [(3000, n, 2000, 2 * B + 100)], np.uint8)
    t = c.truth()[0]
    ref = None
    for kernel in (0, 1):
        for epi in (1, 3):
            d, i, cur = c.run(kernel, epi, curves=True)
            assert np.abs(cur - t).max() <= 3e-6 and i[0] == int(t.argmin()) == 3500 - 2000
            ref = ref or (d, i, cur)
            assert np.array_equal(ref[2], cur) and ref[0][0] == d[0] and ref[1][0] == i[0]


def test_emulated_third_body_overflowing_record_slots(emu):
    """A periodic stream: the template matches exactly every 1000 lags, so a lag block holds sixteen runs of equal
    minimum -- more than a CTA's eight record slots.  Eight runs leave as records (k_finish_runs), the others are
    finished inside the match kernel; both routes, and the all-in-kernel mode, must return what every other variant
    returns."""
    period = programme(1000, 77)
    img = np.tile(period, 3 * B // 1000 + 2)[:3 * B]
    src = img.copy()
    queries = [(2000, 3000, 500, 2 * B + 300),        # exact copies at lags 1000, 2000, ...: the first one inside the range wins
               (2345, 700, 0, B)]                      # lag 345, then every 1000
    c = Case(emu, img, src, queries, np.uint8)
    ref = None
    for kernel in (0, 1):
        for epi, records in ((1, False), (3, False), (3, True)):
            d, i, _ = c.run(kernel, epi, curves=False, records=records)
            # every copy is a minimum up to fp32 FFT rounding (~1e-7): which one wins is decided by the exact path,
            # so all variants must agree bit for bit
            assert i[0] % 1000 == 500 and i[1] % 1000 == 345 and float(d.max()) <= 1e-6, (kernel, epi, records, i, d)
            ref = ref or (d, i)
            assert np.array_equal(ref[0], d) and np.array_equal(ref[1], i), (kernel, epi, records)
            if records:
                assert c.last_record_counts.max() == emu.emu_run_slots()        # the slots did overflow


def test_emulated_persistent_pair_kernel_walks_several_pairs(emu, case_u8):
    """k_match_pair with body 3 is persistent: with fewer CTAs than pairs every CTA walks several pairs -- the barrier's
    phase counter, the record counter, the prefetched descriptor and the tensor-memory columns carry over from one pair
    to the next.  Same answers as one CTA per pair, bit for bit, curves and records included (the first-version
    instantiations keep one CTA per pair: the grid setting does not touch them)."""
    c = case_u8
    want = {(epi, curves): c.run(1, epi, curves) for epi in (1, 3) for curves in (False, True)}
    try:
        for grid in (1, 2, 3):
            emu.emu_set_pair_grid(grid)
            for (epi, curves), w in want.items():
                d, i, cur = c.run(1, epi, curves)
                assert np.array_equal(d, w[0]) and np.array_equal(i, w[1]), (grid, epi, curves)
                if curves:
                    assert np.array_equal(cur, w[2])
    finally:
        emu.emu_set_pair_grid(0)
    n_img = 4 * B - 3000
    rng = np.random.default_rng(9)
    img = (programme(n_img, 3).astype(np.float32) / 255.0).astype(np.float32)
    src = (np.roll(img, -300) + rng.normal(0, 0.01, n_img)).astype(np.float32)
    cf = Case(emu, img, src, [(20000, 18000, 5, 2 * B + 5000), (100, 9000, B, 2 * B), (5000, 40000, 0, 3 * B - 44000)], np.float32)
    w = cf.run(1, 1, curves=False)
    try:
        emu.emu_set_pair_grid(2)
        d, i, _ = cf.run(1, 1, curves=False)
    finally:
        emu.emu_set_pair_grid(0)
    assert np.array_equal(d, w[0]) and np.array_equal(i, w[1])
