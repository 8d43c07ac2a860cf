"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and refuses to run without a GPU (no silent CPU fallback)."""
import ctypes
import os
import re

import pytest

from sushi_b200 import _native
from sushi_b200.common import SushiError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'sushi_b200.h')


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
    syms = header_symbols()
    for must in ('sb_init', 'sb_stream_create', 'sb_find', 'sb_find_batch', 'sb_find_batch_device',
                 'sb_load_pcm', 'sb_normalise', 'sb_last_error'):
        assert must in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_native.LIB_PATH), 'build the library first (__graft_entry__.build())'
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), 'libsushi_b200.so does not export {0}'.format(name)


def test_binding_table_matches_header():
    assert sorted(_native.PROTOTYPES) == header_symbols()


def test_abi_version():
    lib = _native.load_library()
    assert lib.sb_abi_version() == _native.ABI_VERSION
    text = open(HEADER).read()
    assert '#define SB_ABI_VERSION {0}'.format(_native.ABI_VERSION) in text


def _has_gpu():
    try:
        drv = ctypes.CDLL('libcuda.so.1')
        n = ctypes.c_int(0)
        return drv.cuInit(0) == 0 and drv.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


@pytest.mark.skipif(_has_gpu(), reason='only meaningful on a box without a GPU')
def test_no_cpu_fallback_without_gpu():
    lib = _native.load_library()
    rc = lib.sb_init(0)
    assert rc != 0
    assert b'no CPU path' in lib.sb_last_error()
    with pytest.raises(SushiError):
        _native.lib()
    # calls before a successful init fail with a message instead of crashing
    h = ctypes.c_void_p()
    buf = (ctypes.c_uint8 * 16)()
    assert lib.sb_stream_create(buf, 16, 0, ctypes.byref(h)) != 0
    assert b'not initialised' in lib.sb_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'sushi_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'import cv2' not in text, f


def test_missing_library_is_an_error_not_a_fallback(monkeypatch):
    """No .so -> SushiError that says how to build it; nothing falls back to a CPU implementation."""
    monkeypatch.setattr(_native, '_lib', None)
    monkeypatch.setattr(_native, 'LIB_PATH', '/nonexistent/libsushi_b200.so')
    with pytest.raises(SushiError) as e:
        _native.load_library()
    assert 'no CPU fallback' in str(e.value)
