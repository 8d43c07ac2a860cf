"""Executable specification of the index algebra of the packed kernels (sushi_b200/csrc/sb_fused2.cu), in NumPy:
quad-layout chunks -> Hermitian packing + first radix-2 step (pack_quad) -> three radix-16 passes by decimation in
frequency over the (u, v) pairs in the [16][16][32] buffer with strides (529, 33, 1) (fft_passes_dif: pass 1 across the
CTA, passes 2 and 3 inside one warp each, every thread writing over its own inputs) -> last radix-2 step with the
twiddle W32^k3 in the epilogue, under both thread-to-lag mappings (finish_item: 8 lags in each of four rounds;
finish_item_v3: 32 consecutive lags).  The addresses are the kernel's; the result must be the unnormalised inverse
real FFT of the product spectrum, and every access pattern must hit sixteen distinct 8-byte banks per half warp.
CPU only: it guards the algebra, not the CUDA code (the emulation and the GPU parity tests do that)."""
import numpy as np

B, T = 16384, 512
DA, DB = 529, 33


def phys(n):
    return (n >> 9) * DA + ((n >> 5) & 15) * DB + (n & 31)


def pack_quad(a, m, c, s):
    """a = (Y[i], Y[i+B/2]), m = (Y[B-i], Y[B/2-i]) -> chunks C[i] = (u[i], v[i]) and C[B/2-i]."""
    aR, aI, mR, mI = a.real, a.imag, m.real, m.imag
    eR, eI, dR, dI = aR + mR, aI - mI, aR - mR, aI + mI
    wR, wI = np.stack([c, -s], 1), np.stack([s, c], 1)
    oR, oI = dR * wR - dI * wI, dR * wI + dI * wR
    zloR, zloI, zhiR, zhiI = eR - oI, eI + oR, eR + oI, oR - eI
    c2, s2 = c * c - s * s, 2 * c * s
    ur, ui = zloR[:, 0] + zloR[:, 1], zloI[:, 0] + zloI[:, 1]
    dr, di = zloR[:, 0] - zloR[:, 1], zloI[:, 0] - zloI[:, 1]
    lo = (ur + 1j * ui, (dr * c2 - di * s2) + 1j * (dr * s2 + di * c2))
    ur, ui = zhiR[:, 1] + zhiR[:, 0], zhiI[:, 1] + zhiI[:, 0]
    dr, di = zhiR[:, 1] - zhiR[:, 0], zhiI[:, 1] - zhiI[:, 0]
    hi = (ur + 1j * ui, (-(dr * c2) - di * s2) + 1j * (dr * s2 - di * c2))
    return lo, hi


def brev16(r):
    return int('{:04b}'.format(r)[::-1], 2)


def dft16_dif(v):
    v = v.copy()
    h = 8
    while h >= 1:
        for g in range(0, 16, 2 * h):
            for a in range(h):
                x, y = v[g + a].copy(), v[g + a + h].copy()
                v[g + a] = x + y
                v[g + a + h] = (x - y) * np.exp(2j * np.pi * (a * (16 // h)) / 32)
        h //= 2
    return v


def test_packed_pipeline_is_the_inverse_real_fft():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2 * B)
    Y = np.fft.rfft(x)                       # product spectrum, bins 0 .. B
    want = x * 2 * B                         # unnormalised inverse

    i = np.arange(B // 4 + 1)
    A = np.stack([Y[i], Y[i + B // 2]], 1)
    M = np.stack([Y[B - i], Y[B // 2 - i]], 1)
    (lo_u, lo_v), (hi_u, hi_v) = pack_quad(A, M, np.cos(np.pi * i / B), np.sin(np.pi * i / B))

    n_phys = 16 * DA
    bufs = [np.zeros(n_phys, complex), np.zeros(n_phys, complex)]      # u and v halves of every chunk
    for tid in range(T):
        tm = (T - tid) & (T - 1)
        for uu in range(8):
            q = tid + 512 * uu
            for b_, (lo, hi) in zip(bufs, ((lo_u, hi_u), (lo_v, hi_v))):
                b_[phys(tid) + DA * uu] = lo[q]
                if tid != 0:
                    b_[phys(tm) + DA * (15 - uu)] = hi[q]
                elif uu != 0:
                    b_[phys(tm) + DA * (16 - uu)] = hi[q]
    bufs[0][phys(4096)], bufs[1][phys(4096)] = lo_u[4096], lo_v[4096]

    # the chunks are the two half-size sequences u, v of the decimation-in-frequency split
    k = np.arange(B)
    Z = (Y[k] + np.conj(Y[B - k])) + 1j * np.exp(1j * np.pi * k / B) * (Y[k] - np.conj(Y[B - k]))
    u = Z[:B // 2] + Z[B // 2:]
    v = (Z[:B // 2] - Z[B // 2:]) * np.exp(2j * np.pi * np.arange(B // 2) / B)
    c = np.arange(8192)
    pc = np.array([phys(int(n)) for n in c])
    assert np.abs(bufs[0][pc] - u).max() < 1e-9 * np.abs(u).max()
    assert np.abs(bufs[1][pc] - v).max() < 1e-9 * np.abs(v).max()

    tid = np.arange(T)
    lane, warp = tid & 31, tid >> 5
    for buf in bufs:
        at = warp * DB + lane                                                      # pass 1: over a, position t = tid
        vv = dft16_dif(np.stack([buf[at + DA * a] for a in range(16)]))
        for k1 in range(16):
            buf[at + DA * k1] = vv[brev16(k1)] * np.exp(2j * np.pi * tid * k1 / 8192)
        at = warp * DA + lane                                                      # pass 2: warp k1, over b, position l = lane
        vv = dft16_dif(np.stack([buf[at + DB * b_] for b_ in range(16)]))
        for k2 in range(16):
            buf[at + DB * k2] = vv[brev16(k2)] * np.exp(2j * np.pi * lane * k2 / 512)
        at = warp * DA + (lane & 15) * DB + (lane >> 4)                            # pass 3: lane (k2, m), over c
        vv = dft16_dif(np.stack([buf[at + 2 * c_] for c_ in range(16)]))
        for k3 in range(16):
            buf[at + 2 * k3] = vv[brev16(k3)]

    # epilogue, first version: chunk 1024c + 2tid + e, twiddle W32^(tid/128) * W8^c
    out = np.zeros(B)
    for c_ in range(4):
        for e in range(2):
            p = 2 * (tid & 7) * DA + ((tid >> 3) & 15) * DB + 2 * (tid >> 7) + DA * e + 8 * c_
            w = np.exp(2j * np.pi * (tid >> 7) / 32) * np.exp(2j * np.pi * c_ / 8)
            xu = bufs[0][p] + w * bufs[0][p + 1]
            xv = bufs[1][p] + w * bufs[1][p + 1]
            m0 = c_ * 4096 + tid * 8 + 4 * e
            out[m0], out[m0 + 1], out[m0 + 2], out[m0 + 3] = xu.real, xu.imag, xv.real, xv.imag
    assert np.abs(out - want[:B]).max() < 1e-9 * np.abs(want).max()

    # epilogue, body 3: chunks 8*tid + i (32 consecutive lags per thread), twiddle W32^warp
    out3 = np.zeros(B)
    for i_ in range(8):
        p = 8 * (tid & 1) * DA + ((tid >> 1) & 15) * DB + 2 * warp + DA * i_
        w = np.exp(2j * np.pi * warp / 32)
        xu = bufs[0][p] + w * bufs[0][p + 1]
        xv = bufs[1][p] + w * bufs[1][p + 1]
        m0 = 32 * tid + 4 * i_
        out3[m0], out3[m0 + 1], out3[m0 + 2], out3[m0 + 3] = xu.real, xu.imag, xv.real, xv.imag
    assert np.abs(out3 - want[:B]).max() < 1e-9 * np.abs(want).max()

    # the exact path's general address of chunk jj: [jj % 16][(jj / 16) % 16][2 (jj / 256)], twiddle W32^(jj / 256)
    jj = np.arange(4096)
    p = (jj & 15) * DA + ((jj >> 4) & 15) * DB + 2 * (jj >> 8)
    xu = bufs[0][p] + np.exp(2j * np.pi * (jj >> 8) / 32) * bufs[0][p + 1]
    assert np.abs(xu.real - want[0:B:4]).max() < 1e-9 * np.abs(want).max()


def test_buffer_strides_are_bank_conflict_free():
    """Every warp-wide 64-bit access of the passes and of both epilogues: the sixteen lanes of a half warp hit sixteen
    distinct 8-byte banks (addresses are in 8-byte units of the two float2 arrays)."""
    def ok(addr):                            # addr: 32 element indices of one warp-wide access
        return all(len(set(int(a) % 16 for a in addr[h:h + 16])) == 16 for h in (0, 16))
    lane = np.arange(32)
    for warp in range(16):
        tid = warp * 32 + lane
        for r in range(16):
            assert ok(warp * DB + lane + DA * r)                                        # pass 1 loads / stores; packing stores
            assert ok(warp * DA + lane + DB * r)                                        # pass 2
            assert ok(warp * DA + (lane & 15) * DB + (lane >> 4) + 2 * r)               # pass 3
        for c_ in range(4):
            for e in range(2):
                p = 2 * (tid & 7) * DA + ((tid >> 3) & 15) * DB + 2 * (tid >> 7) + DA * e + 8 * c_
                assert ok(p) and ok(p + 1)                                              # epilogue, first version
        for i_ in range(8):
            p = 8 * (tid & 1) * DA + ((tid >> 1) & 15) * DB + 2 * warp + DA * i_
            assert ok(p) and ok(p + 1)                                                  # epilogue, body 3
    # mirrored chunks of the packing stage: thread tid writes C[B/2 - i] into column (512 - tid) % 512 -- descending
    # addresses whose first lane sits in the next row of the buffer: one pair of lanes per warp shares a bank (2-way
    # on one of the 32 stores' half warps), nothing worse
    for warp in range(16):
        tid = warp * 32 + lane
        tm = (T - tid) & (T - 1)
        col = np.array([phys(int(t)) for t in tm])
        for h in (0, 16):
            banks = [int(a) % 16 for a in col[h:h + 16]]
            assert max(banks.count(b_) for b_ in set(banks)) <= 2


def test_quad_row_layout_is_a_bijection():
    """qa()/qm(): blocks of 256 A chunks followed by 256 M chunks, the self-mirrored quad behind them."""
    q4 = B // 4

    def qa(i):
        return (i >> 8) * 512 + (i & 255) if i < q4 else 2 * q4

    def qm(i):
        return qa(i) + 256 if i < q4 else 2 * q4 + 1
    slots = [qa(i) for i in range(q4 + 1)] + [qm(i) for i in range(q4 + 1)]
    assert sorted(slots) == list(range(2 * q4 + 2)) and 2 * q4 + 2 <= 8200
    # kernel A's per-thread addressing: quad tid + 512*uu sits at (tid>>8)*512 + (tid&255) + 1024*uu
    for tid in (0, 1, 255, 256, 511):
        for uu in range(8):
            assert qa(tid + 512 * uu) == (tid >> 8) * 512 + (tid & 255) + 1024 * uu
