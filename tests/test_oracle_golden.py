"""The oracle (oracle/ref_*.py, a restatement) against the golden vectors produced by THE
REFERENCE ITSELF in this image (oracle/gen_golden.py).  Same host + same cv2 => these match
exactly; the slack covers cv2's ISA dispatch on a different host."""
import zlib

import numpy as np
import pytest

from oracle import ref_loader, ref_matcher
from tests.helpers import oracle_stream_from_pcm

LOADER_CASES = ['mono12k', 'stereo48k', 'mono44k1', 'stereo22k05', 'mono8k_up', 'six48k']


@pytest.mark.parametrize('name', LOADER_CASES)
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_loader_restatement_matches_reference(golden_loader, name, stype):
    g = golden_loader
    fr, ch, sr = [int(v) for v in g[name + '_spec']]
    s = oracle_stream_from_pcm(g[name + '_pcm'], fr, ch, sr, stype)
    ref = g['{0}_{1}_data'.format(name, stype)]
    rate, count, pad = [int(v) for v in g['{0}_{1}_meta'.format(name, stype)]]
    assert (s.sample_rate, int(s.sample_count), s.padding_size) == (rate, count, pad)
    assert s.data.dtype == ref.dtype and s.data.shape == ref.shape
    assert np.array_equal(s.data, ref)


LOADER24_CASES = ['stereo48k_24', 'mono44k1_24', 'six48k_24']


@pytest.mark.parametrize('name', LOADER24_CASES)
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_loader_restatement_matches_reference_int24(golden_loader24, tmp_path, name, stype):
    """The int24 branch (wav.py:71-74), pinned since round 2: the oracle's load_wav on the frozen 24-bit files
    against WavStream.data of the reference itself, bit for bit."""
    g = golden_loader24
    p = str(tmp_path / 'x.wav')
    open(p, 'wb').write(g[name + '_wav'].tobytes())
    data, count, pad = ref_loader.load_wav(p, 12000, stype)
    ref = g['{0}_{1}_data'.format(name, stype)]
    rate, rcount, rpad = [int(v) for v in g['{0}_{1}_meta'.format(name, stype)]]
    assert (int(count), pad) == (rcount, rpad) and rate == 12000
    assert data.dtype == ref.dtype and data.shape == ref.shape and np.array_equal(data, ref)


@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_matcher_restatement_matches_reference(golden_matcher, stype):
    g = golden_matcher
    src = oracle_stream_from_pcm(g['src_pcm'], 12000, 1, 12000, stype)
    dst = oracle_stream_from_pcm(g['dst_pcm'], 12000, 1, 12000, stype)
    assert zlib.crc32(src.data.tobytes()) == int(g['src_{0}_crc'.format(stype)][0])
    assert zlib.crc32(dst.data.tobytes()) == int(g['dst_{0}_crc'.format(stype)][0])
    for q, (a, b, c, w) in enumerate(g['queries']):
        d, t = dst.find_substream(src.get_substream(a, b), c, w)
        assert isinstance(d, np.float32)
        assert abs(float(d) - float(g['diff_' + stype][q])) <= 5e-6, q
        assert abs(t - g['time_' + stype][q]) <= 1.0 / 12000 + 1e-12, q


def test_known_answers(golden_matcher):
    """Degenerate inputs: the values cv2 gives in this image."""
    g = golden_matcher
    assert np.all(g['deg_zero_window'] == 1.0)
    assert np.all(g['deg_zero_template'] == 1.0)
    assert np.allclose(g['deg_const_7_vs_9'], 0.063492, atol=1e-6)
    assert g['deg_periodic'].argmin() == 0           # first of equal minima
    # and the fp64 closed form agrees with cv2 on them
    nine = np.full(64, 9.0)
    assert np.allclose(ref_matcher.sqdiff_normed_fp64(nine, np.full(8, 7.0)), g['deg_const_7_vs_9'], atol=1e-6)


@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_fp64_truth_brackets_cv2(golden_matcher, stype):
    """cv2 stays within ~1e-6 of the closed form: the arbitration tool is sound."""
    g = golden_matcher
    src = oracle_stream_from_pcm(g['src_pcm'], 12000, 1, 12000, stype)
    dst = oracle_stream_from_pcm(g['dst_pcm'], 12000, 1, 12000, stype)
    s0, s1, stride = [int(v) for v in g['curve0_{0}_s0'.format(stype)]]
    a, b, c, w = g['queries'][0]
    pat = src.get_substream(a, b)
    truth = ref_matcher.sqdiff_normed_fp64(dst.data[0, s0:s1], pat[0])
    assert np.abs(truth - g['curve0_' + stype]).max() < 3e-6
    assert truth.argmin() == g['curve0_' + stype].argmin()
