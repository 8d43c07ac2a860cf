"""Executable specification of the index algebra of the packed kernels (sushi_b200/csrc/sb_fused2.cu), in NumPy:
quad-layout chunks -> Hermitian packing + first radix-2 step (pack_quad) -> three radix-16 Stockham passes over
the (u, v) pairs in the padded buffer (fft_passes) -> last radix-2 step in the epilogue (finish_item).  The
addresses (phys(), the 544 / 17 / 272 / 4352 / 1088 strides, the mirrored-chunk rule) are the kernel's; the
result must be the unnormalised inverse real FFT of the product spectrum.  CPU only: it guards the algebra,
not the CUDA code (the GPU parity tests do that)."""
import numpy as np

B, T = 16384, 512


def phys(c):
    return c + (c >> 4)


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

    n_phys = 8192 + 512
    bufs = [np.zeros(n_phys, complex), np.zeros(n_phys, complex)]      # u and v halves of every chunk
    for tid in range(T):
        tm = (T - tid) & (T - 1)
        for uu in range(8):
            q = tid + 512 * uu
            for b_, (lo, hi) in zip(bufs, ((lo_u, hi_u), (lo_v, hi_v))):
                b_[phys(tid) + 544 * uu] = lo[q]
                if tid != 0:
                    b_[phys(tm) + 544 * (15 - uu)] = hi[q]
                elif uu != 0:
                    b_[phys(tm) + 544 * (16 - uu)] = hi[q]
    bufs[0][phys(4096)], bufs[1][phys(4096)] = lo_u[4096], lo_v[4096]

    # the chunks are the two half-size sequences u, v of the decimation-in-frequency split
    k = np.arange(B)
    Z = (Y[k] + np.conj(Y[B - k])) + 1j * np.exp(1j * np.pi * k / B) * (Y[k] - np.conj(Y[B - k]))
    u = Z[:B // 2] + Z[B // 2:]
    v = (Z[:B // 2] - Z[B // 2:]) * np.exp(2j * np.pi * np.arange(B // 2) / B)
    c = np.arange(8192)
    assert np.abs(bufs[0][phys(c)] - u).max() < 1e-9 * np.abs(u).max()
    assert np.abs(bufs[1][phys(c)] - v).max() < 1e-9 * np.abs(v).max()

    tid = np.arange(T)
    for buf in bufs:
        src = phys(tid)
        vv = dft16_dif(np.stack([buf[src + 544 * r] for r in range(16)]))          # pass 1
        for r in range(16):
            buf[17 * tid + r] = vv[brev16(r)]
        kk = tid & 15                                                              # pass 2
        vv = np.stack([buf[src + 544 * r] for r in range(16)])
        for r in range(1, 16):
            vv[r] = vv[r] * np.exp(2j * np.pi * r * kk / 256)
        vv = dft16_dif(vv)
        for r in range(16):
            buf[272 * (tid >> 4) + kk + 17 * r] = vv[brev16(r)]
        kk = tid & 255                                                             # pass 3
        vv = np.stack([buf[src + 544 * r] for r in range(16)])
        for r in range(1, 16):
            vv[r] = vv[r] * np.exp(2j * np.pi * r * kk / 4096)
        vv = dft16_dif(vv)
        for r in range(16):
            buf[4352 * (tid >> 8) + phys(kk) + 272 * r] = vv[brev16(r)]

    out = np.zeros(B)
    for c_ in range(4):                                                            # epilogue rounds
        for e in range(2):
            p = 2 * tid + (tid >> 3) + 1088 * c_ + e
            w = np.exp(2j * np.pi * (2 * tid + e) / 8192) * np.exp(2j * np.pi * c_ / 8)
            xu = bufs[0][p] + w * bufs[0][p + 4352]
            xv = bufs[1][p] + w * bufs[1][p + 4352]
            m0 = c_ * 4096 + tid * 8 + 4 * e
            out[m0], out[m0 + 1], out[m0 + 2], out[m0 + 3] = xu.real, xu.imag, xv.real, xv.imag
    assert np.abs(out - want[:B]).max() < 1e-9 * np.abs(want).max()


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
