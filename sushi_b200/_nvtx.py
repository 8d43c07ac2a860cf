"""Optional NVTX ranges around the host-side phases (tracing).

Off unless SUSHI_B200_NVTX=1: then libnvToolsExt is opened with ctypes (the CUDA toolkit's or the one the Python
environment ships) and `with nvtx_range('name'):` brackets a phase for Nsight Systems / ncu --nvtx.  Without the
library, or with the variable unset, the context manager does nothing -- per-kernel device times come from the
library's own sb_profile_* brackets either way."""
import contextlib
import ctypes
import glob
import os
import sys

_lib = None
_tried = False


def _load():
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    if os.environ.get('SUSHI_B200_NVTX', '0') != '1':
        return None
    names = ['libnvToolsExt.so.1', 'libnvToolsExt.so']
    dirs = ['', os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'lib64')]
    for p in sys.path:
        dirs += glob.glob(os.path.join(p, 'nvidia', 'nvtx', 'lib'))
    for d in dirs:
        for n in names:
            try:
                lib = ctypes.CDLL(os.path.join(d, n) if d else n)
                lib.nvtxRangePushA.argtypes = [ctypes.c_char_p]
                lib.nvtxRangePushA.restype = ctypes.c_int
                lib.nvtxRangePop.restype = ctypes.c_int
                _lib = lib
                return _lib
            except (OSError, AttributeError):
                continue
    return None


def enabled():
    return _load() is not None


@contextlib.contextmanager
def nvtx_range(name):
    lib = _load()
    if lib is None:
        yield
        return
    lib.nvtxRangePushA(name.encode('ascii', 'replace'))
    try:
        yield
    finally:
        lib.nvtxRangePop()
