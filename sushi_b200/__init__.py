"""sushi_b200 -- B200-native implementation of tp7/Sushi's audio template-matching path.

Public surface (mirrors the reference's wav.py / the part of sushi.py that drives it):
    WavStream            drop-in stream class, GPU-backed find_substream
    SushiError, clip     as in the reference's common.py
"""
from .common import SushiError, clip, format_time   # noqa: F401
from .wavstream import WavStream, DownmixedWavFile   # noqa: F401

__version__ = '0.1.0'
