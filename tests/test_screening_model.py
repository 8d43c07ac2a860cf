"""NumPy restatement of the two fp32 screening loops of the packed kernels (sb_fused2.cu, finish_item) against
the exact fp64 value (sb_fused_common.cuh, sqdiff_exact).  Screening only has to rank the lags of a block to
within half the screening margin; the lags inside the margin are then evaluated exactly.  This file pins the
algebra of the trimmed per-lag loop (body 3 runs it on the runs its bounds select; as the loop over all lags it was variant 2, dropped in round 2):

    v' = (A + rq - 2b*rs - 2*scale*cc) * rsqrt(wq + 0.25),   A = w0q + sum T^2 - 2(b*w0s + k)   (one fp64 rounding)

is value * sqrt(sum T^2) for unsaturated lags, stays finite on silent windows, and never falls below the exact
value by more than the margin, which is what the candidate test needs.  float32 arithmetic is emulated
step by step (every operation rounded to float32, fused multiply-adds through float64 and one rounding)."""
import numpy as np
import pytest

F = np.float32
MARGIN = 8e-6           # kScreenMargin
B = 16384


def fma32(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def exact_value(cc, wsum, wsq, a, b, tsum, tsq, n):
    """sqdiff_exact: OpenCV's rule, sum(I*T) kept as float32."""
    sit = cc + b * wsum + a * tsum - n * a * b
    corr = np.float64(F(sit))
    num = max(wsq - 2.0 * corr + tsq, 0.0)
    p = wsq * tsq
    if not wsq > 0.0 or not p > 0.0:
        return F(1.0)
    t = np.sqrt(p)
    return F(num / t) if num < t else F(1.0)


def screen_both(img, tmpl, j0):
    """Run of 8 lags starting at j0: (exact values, v1 screening values, v2 screening values / sqrt(sum T^2))."""
    n = tmpl.size
    I = img.astype(np.float64)
    T = tmpl.astype(np.float64)
    a = np.rint(I.mean())
    tsum, tsq = T.sum(), (T * T).sum()
    b = np.rint(tsum / n)                         # Acc<uint8>::centre
    k_const = a * tsum - n * a * b
    scale = 1.0 / (2 * B)
    w0s, w0q = I[j0:j0 + n].sum(), (I[j0:j0 + n] ** 2).sum()
    f_tsq, f_b, f_scale = F(tsq), F(b), F(scale)
    # v1 run constants
    f_w0q1, f_k0 = F(w0q), F(b * w0s + k_const)
    # v2 run constants
    f_w0q2, f_A = F(w0q + 0.25), F(w0q + tsq - 2.0 * (b * w0s + k_const))
    m2s, m2b = F(-2.0) * f_scale, F(-2.0) * f_b
    rt = np.sqrt(F(tsq))
    ex, v1s, v2s = [], [], []
    rq = rs = 0
    for i in range(8):
        j = j0 + i
        w = I[j:j + n]
        corr_c = float(((T - b) * (w - a)).sum())                 # what the inverse FFT delivers, times 1/(2B)
        cc = F(corr_c * 2 * B)                                     # fp32 FFT output (unscaled)
        ex.append(exact_value(np.float64(cc) * scale, w.sum(), (w * w).sum(), a, b, tsum, tsq, n))
        # ---- v1
        wq = F(f_w0q1 + F(rq))
        sit = fma32(cc, f_scale, fma32(f_b, F(rs), f_k0))
        num = max(F(F(wq + f_tsq) - F(F(2.0) * sit)), F(0.0))
        pr = F(wq * f_tsq)
        with np.errstate(divide='ignore', invalid='ignore'):
            v = F(num * F(1.0 / np.sqrt(np.float64(pr)))) if pr > 0 else F(np.inf)
        v1s.append(F(1.0) if not v < 1.0 else v)                   # fminf(NaN/inf, 1) = 1
        # ---- v2
        frq = F(rq)
        num2 = fma32(cc, m2s, fma32(m2b, F(rs), F(f_A + frq)))
        v2 = F(num2 * F(1.0 / np.sqrt(np.float64(F(f_w0q2 + frq)))))
        v2s.append(v2 / rt if rt > 0 else np.inf)
        lo, hi = int(img[j]), int(img[j + n]) if j + n < img.size else 0
        rq += (hi + lo) * (hi - lo)
        rs += hi - lo
    return np.array(ex, np.float64), np.array(v1s, np.float64), np.array(v2s, np.float64)


@pytest.mark.parametrize('seed', range(6))
def test_trimmed_screening_tracks_the_exact_value(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(600, 40000))
    img = rng.integers(0, 256, n + 4000, dtype=np.uint8)
    if seed % 2:                                   # programme-like: clipped rails, a correlated template
        img = np.clip(np.rint(128 + 90 * np.convolve(rng.standard_normal(img.size + 8), np.hanning(9), 'valid')), 0, 255).astype(np.uint8)
    j_true = int(rng.integers(100, 3000))
    tmpl = np.clip(img[j_true:j_true + n].astype(np.int32) + rng.integers(-6, 7, n), 0, 255).astype(np.uint8)
    worst1 = worst2 = 0.0
    for j0 in (j_true - 3, j_true - 8, 0, 17, 3500):
        ex, v1, v2 = screen_both(img, tmpl, j0)
        unsat = ex < 1.0
        worst1 = max(worst1, np.abs(v1 - ex)[unsat].max(initial=0.0))
        worst2 = max(worst2, np.abs(v2 - ex)[unsat].max(initial=0.0))
        assert np.all(v2 >= ex - MARGIN / 2)        # saturated lags only ever screen HIGHER than their value
    assert worst1 <= MARGIN / 2 and worst2 <= MARGIN / 2, (worst1, worst2)


def test_trimmed_screening_on_silence_and_zero_template():
    n = 900
    img = np.zeros(4000, np.uint8)
    img[2000:] = np.random.default_rng(1).integers(0, 256, 2000)
    tmpl = np.random.default_rng(2).integers(1, 256, n).astype(np.uint8)
    ex, v1, v2 = screen_both(img, tmpl, 100)       # windows of pure silence: exact value 1, screening finite and >= 1
    assert np.all(ex == 1.0) and np.all(np.isfinite(v2)) and np.all(v2 >= 1.0 - MARGIN)
    ex, v1, v2 = screen_both(img, tmpl, 1096)      # the window slides out of the silence inside the run
    assert np.all(np.isfinite(v2)) and np.all(v2 >= ex - MARGIN / 2)
    ex, v1, v2 = screen_both(img, np.zeros(n, np.uint8), 2100)    # zero template: sum T^2 = 0, everything saturates
    assert np.all(ex == 1.0) and np.all(v2 == np.inf)             # -> the kernel's "evaluate every lag" fallback
