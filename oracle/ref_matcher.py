"""CPU restatement of the reference matcher -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/wav.py:164-188 (duration_seconds, get_substream,
_get_sample_for_time, find_substream) and common.py:41-42 (clip).  The per-lag
arithmetic is OpenCV's: cv2.matchTemplate(..., TM_SQDIFF_NORMED), wav.py:185.
"""
import cv2
import numpy as np

PADDING_SECONDS = 10          # wav.py:106


def clip(value, minimum, maximum):            # common.py:41-42
    return max(min(value, maximum), minimum)


class RefStream(object):
    """What a loaded reference WavStream looks like to its callers: data (1,N),
    sample_rate, sample_count, padding_size (wav.py:116-120)."""

    def __init__(self, data, sample_rate, padding_size, sample_count):
        self.data = data
        self.sample_rate = sample_rate
        self.padding_size = padding_size
        self.sample_count = sample_count

    @property
    def duration_seconds(self):                 # wav.py:164-166
        return self.sample_count / self.sample_rate

    def sample_for_time(self, timestamp):       # wav.py:173-175
        return int(self.sample_rate * timestamp) + self.padding_size

    def get_substream(self, start, end):        # wav.py:168-171
        return self.data[:, self.sample_for_time(start):self.sample_for_time(end)]

    def find_substream(self, pattern, window_center, window_size):     # wav.py:177-188
        start_time = clip(window_center - window_size, -PADDING_SECONDS, self.duration_seconds)
        end_time = clip(window_center + window_size, 0, self.duration_seconds + PADDING_SECONDS)
        start_sample = self.sample_for_time(start_time)
        end_sample = self.sample_for_time(end_time) + len(pattern[0])
        search_source = self.data[:, start_sample:end_sample]
        result = cv2.matchTemplate(search_source, pattern, cv2.TM_SQDIFF_NORMED)
        min_idx = result.argmin(axis=1)[0]
        return result[0][min_idx], start_time + (min_idx / float(self.sample_rate))

    def match_curve(self, pattern, start_sample, nlags):
        """The whole curve of one query given integer offsets (for curve-level parity)."""
        n = len(pattern[0])
        return cv2.matchTemplate(self.data[:, start_sample:start_sample + nlags + n - 1], pattern,
                                 cv2.TM_SQDIFF_NORMED)[0]


def sqdiff_normed_fp64(image, templ):
    """Closed-form fp64 'truth' of TM_SQDIFF_NORMED on 1-D inputs, used to arbitrate when
    the GPU and cv2 disagree: max(sum I^2 - 2 sum IT + sum T^2, 0) / (sqrt(sum I^2) sqrt(sum T^2)),
    1.0 where the quotient is not < 1 (OpenCV's saturation rule)."""
    image = np.asarray(image, np.float64).ravel()
    templ = np.asarray(templ, np.float64).ravel()
    n = templ.size
    nl = image.size - n + 1
    size = 1 << int(np.ceil(np.log2(image.size + n)))
    corr = np.fft.irfft(np.fft.rfft(image, size) * np.conj(np.fft.rfft(templ, size)), size)[:nl]
    psq = np.concatenate([[0.0], np.cumsum(image * image)])
    wnd = psq[n:n + nl] - psq[:nl]
    tsq = float(np.dot(templ, templ))
    num = np.maximum(wnd - 2.0 * corr + tsq, 0.0)
    t = np.sqrt(np.maximum(wnd, 0.0)) * np.sqrt(tsq)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.where(np.abs(num) < t, num / t, 1.0)
    return out
