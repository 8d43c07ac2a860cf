"""Host-memory hygiene for long-running drivers (bench.py, the test-suite).

glibc returns large freed blocks to the kernel and maps fresh pages for every big numpy
temporary; on virtualised hosts the first touch of a fresh page can be very slow.  Keeping the
heap (no mmap for big blocks, no trimming) makes repeated multi-hundred-MB temporaries cheap.
"""
import ctypes

M_TRIM_THRESHOLD, M_TOP_PAD, M_MMAP_THRESHOLD, M_MMAP_MAX = -1, -2, -3, -4


def keep_heap(top_pad=64 << 20):
    try:
        libc = ctypes.CDLL('libc.so.6')
        libc.mallopt(M_MMAP_MAX, 0)
        libc.mallopt(M_TRIM_THRESHOLD, 0x7fffffff)
        libc.mallopt(M_TOP_PAD, int(top_pad))
        return True
    except Exception:      # non-glibc platform: nothing to tune
        return False
