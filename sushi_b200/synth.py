"""Seeded synthetic inputs for the benchmark configurations (SURVEY.md section 8d / BASELINE.json
`configs`): 12 kHz mono int16 PCM that looks like programme audio (aperiodic low-passed noise under a
piecewise-constant loudness envelope), a source stream that is the destination moved by a known
shift function plus independent noise, and non-nested subtitle event lists.  No file or network
access; everything derives from numpy.random.default_rng(seed).  Generation is chunked and
in place so that a 90-minute pair needs little more host memory than its int16 output.
"""
import numpy as np

_CHUNK = 1 << 20
_KERNEL = np.hanning(9).astype(np.float32)
_KERNEL /= _KERNEL.sum()
_SIGMA = float(np.sqrt((_KERNEL.astype(np.float64) ** 2).sum()))


def programme_audio(n, seed, rate=12000, amplitude=6000.0):
    """int16 array of n samples: low-passed Gaussian noise x loudness steps of 0.2 s, peak ~ +-amplitude."""
    rng = np.random.default_rng(seed)
    step = int(0.2 * rate)
    gains = rng.uniform(0.05, 1.0, n // step + 2).astype(np.float32) ** 2
    out = np.empty(n, np.int16)
    scale = np.float32(amplitude / (4.0 * _SIGMA))
    tail = rng.standard_normal(8, dtype=np.float32)
    for a in range(0, n, _CHUNK):
        b = min(a + _CHUNK, n)
        x = np.concatenate([tail, rng.standard_normal(b - a, dtype=np.float32)])
        tail = x[-8:].copy()
        y = np.convolve(x, _KERNEL, mode='valid')          # b - a samples
        g = gains[(np.arange(a, b) // step)]
        y *= g
        y *= scale
        np.rint(y, out=y)
        np.clip(y, -32768, 32767, out=y)
        out[a:b] = y
    return out


def make_pair(duration_s, seed, shift_fn=None, rate=12000, noise=0.02, amplitude=6000.0):
    """(src_pcm int16, dst_pcm int16): src(t) = dst(t + shift(t)) + noise.

    shift_fn: None / float -> constant shift in seconds; or a list of (t_start, shift_s) segments
    (piecewise-constant per chapter, BASELINE config 4)."""
    n = int(round(duration_s * rate))
    dst = programme_audio(n, seed, rate, amplitude)
    src = np.zeros(n, np.int16)
    if shift_fn is None:
        shift_fn = 0.0
    segments = [(0.0, float(shift_fn))] if np.isscalar(shift_fn) else list(shift_fn)
    for i, (t0, sh) in enumerate(segments):
        a = int(round(t0 * rate))
        b = int(round(segments[i + 1][0] * rate)) if i + 1 < len(segments) else n
        d = int(round(sh * rate))
        lo, hi = max(a, -d, 0), min(b, n - d, n)
        if hi > lo:
            src[lo:hi] = dst[lo + d:hi + d]
    rng = np.random.default_rng(seed + 7919)
    sigma = np.float32(noise * amplitude)
    for a in range(0, n, _CHUNK):
        b = min(a + _CHUNK, n)
        z = rng.standard_normal(b - a, dtype=np.float32)
        z *= sigma
        z += src[a:b]
        np.rint(z, out=z)
        np.clip(z, -32768, 32767, out=z)
        src[a:b] = z
    return src, dst


def make_events(count, duration_s, seed, min_len=1.0, max_len=4.0, margin=2.0):
    """`count` events (start, end) in seconds on the ASS centisecond grid, sorted, non-nested
    (starts and ends both strictly increase) and each longer than the 0.417 s typesetting
    threshold (sushi.py:759), so that search groups == events."""
    rng = np.random.default_rng(seed + 104729)
    lens = rng.uniform(min_len, max_len, count)
    span = duration_s - 2 * margin - max_len
    if span <= count * 0.02:
        raise ValueError('too many events for this duration')
    starts = np.sort(rng.uniform(margin, margin + span, count))
    starts = np.round(starts * 100).astype(np.int64)
    ends = starts + np.round(lens * 100).astype(np.int64)
    for i in range(1, count):                 # strictly increasing on the centisecond grid
        if starts[i] <= starts[i - 1]:
            starts[i] = starts[i - 1] + 1
        if ends[i] <= ends[i - 1]:
            ends[i] = ends[i - 1] + 1
        if ends[i] < starts[i] + 50:
            ends[i] = starts[i] + 50
    return starts / 100.0, ends / 100.0
