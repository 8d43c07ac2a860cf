"""Host side of the product (no GPU): the loader mirror and the integer query planning must agree
bit for bit with the reference's golden outputs / the oracle restatement."""
import numpy as np
import pytest

from sushi_b200 import wavstream
from sushi_b200.wavstream import WavStream
from tests.helpers import oracle_stream_from_pcm
from tests.test_oracle_golden import LOADER_CASES


def host_stream_from_pcm(pcm, framerate, channels, sample_rate, sample_type):
    """Product loader without the upload step (which needs the GPU)."""
    class _Mem(object):
        pass
    pcm = np.ascontiguousarray(pcm, '<i2')
    raw = pcm.reshape(-1).view(np.uint8)
    mem = _Mem()
    mem.framerate, mem.channels_count, mem.sample_width = framerate, channels, 2
    mem.frame_size = 2 * channels
    mem.frames_count = raw.size // mem.frame_size
    pos = [0]

    def readframes(count):
        a = pos[0]
        b = min(a + count * mem.frame_size, raw.size)
        pos[0] = b
        return wavstream.decode_downmix(raw[a:b].tobytes(), 2, channels)
    mem.readframes = readframes
    s = object.__new__(WavStream)
    s._handle = None
    s._load(mem, sample_rate, sample_type)
    return s


@pytest.mark.parametrize('name', LOADER_CASES)
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_host_loader_matches_reference_golden(golden_loader, name, stype):
    g = golden_loader
    fr, ch, sr = [int(v) for v in g[name + '_spec']]
    s = host_stream_from_pcm(g[name + '_pcm'], fr, ch, sr, stype)
    ref = g['{0}_{1}_data'.format(name, stype)]
    rate, count, pad = [int(v) for v in g['{0}_{1}_meta'.format(name, stype)]]
    assert (s.sample_rate, int(s.sample_count), s.padding_size) == (rate, count, pad)
    assert s.data.dtype == ref.dtype and s.data.shape == ref.shape
    assert np.array_equal(s.data, ref)


@pytest.mark.parametrize('n_in,n_out', [(48000, 12000), (44100, 12000), (22050, 12000), (8000, 12000),
                                        (48000, 11999), (4410, 1200), (12345, 3359), (1, 1), (7, 3)])
def test_nearest_index_map_is_cv2_resize(n_in, n_out):
    import cv2   # the checker, not the product
    row = np.arange(n_in, dtype=np.float32).reshape(1, -1)
    want = cv2.resize(row, (n_out, 1), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
    assert np.array_equal(wavstream.nearest_index_map(n_in, n_out), want)


@pytest.mark.parametrize('name', ['stereo48k_24', 'mono44k1_24', 'six48k_24'])
@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_host_loader_matches_reference_golden_int24(golden_loader24, tmp_path, name, stype):
    """The product's RIFF walk + readframes + NumPy loader mirror on the frozen 24-bit files against the reference's
    own output (no GPU: the upload step is skipped)."""
    g = golden_loader24
    p = str(tmp_path / 'x.wav')
    open(p, 'wb').write(g[name + '_wav'].tobytes())
    f = wavstream.DownmixedWavFile(p)
    try:
        assert f.sample_width == 3
        s = object.__new__(WavStream)
        s._handle = None
        s._load(f, 12000, stype)
    finally:
        f.close()
    ref = g['{0}_{1}_data'.format(name, stype)]
    rate, count, pad = [int(v) for v in g['{0}_{1}_meta'.format(name, stype)]]
    assert (s.sample_rate, int(s.sample_count), s.padding_size) == (rate, count, pad)
    assert s.data.dtype == ref.dtype and np.array_equal(s.data, ref)


def test_int24_decode_takes_top_16_bits():
    vals = np.array([0x123456, -0x123456, 0x7FFFFF, -0x800000, 255, -256], np.int32)
    raw = b''.join(int(v & 0xFFFFFF).to_bytes(3, 'little') for v in vals)
    got = wavstream.decode_downmix(raw, 3, 1)
    assert np.array_equal(got, (vals >> 8).astype(np.float32))


def test_downmix_sums_left_to_right_in_float32():
    pcm = np.array([[30000, 30000, -5], [1, 2, 3]], np.int16)
    got = wavstream.decode_downmix(pcm.tobytes(), 2, 3)
    want = ((np.float32(30000) + np.float32(30000)) + np.float32(-5)) / np.float32(3)
    assert got[0] == want and got.dtype == np.float32


def test_riff_reader(tmp_path):
    import wave
    p = str(tmp_path / 'a.wav')
    pcm = (np.arange(2400 * 2, dtype=np.int16) - 1200).reshape(-1, 2)
    with wave.open(p, 'wb') as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(24000); w.writeframes(pcm.tobytes())
    f = wavstream.DownmixedWavFile(p)
    assert (f.channels_count, f.framerate, f.sample_width, f.frames_count) == (2, 24000, 2, 2400)
    mono = f.readframes(100)
    assert np.array_equal(mono, (pcm[:100, 0].astype(np.float32) + pcm[:100, 1].astype(np.float32)) / np.float32(2))
    f.close()
    bad = str(tmp_path / 'b.wav')
    open(bad, 'wb').write(b'JUNKxxxxWAVE')
    with pytest.raises(wavstream.SushiError):
        wavstream.DownmixedWavFile(bad)


def test_unknown_sample_type_raises():
    with pytest.raises(wavstream.SushiError):
        WavStream('/nonexistent.wav', sample_type='int16')


@pytest.mark.parametrize('stype', ['uint8', 'float32'])
def test_query_planning_matches_reference_calls(golden_shifts, stype):
    """Every find_substream call the reference made in the golden scenarios: our integer planning
    (template offset/length, first lag, lag count, start time) equals the oracle's slicing."""
    from sushi_b200 import synth
    g = golden_shifts
    for name in ('const', 'jump', 'rewind'):
        dur, seed, count = g[name + '_gen']
        shift = [tuple(r) for r in g[name + '_shift']]
        src_pcm, dst_pcm = synth.make_pair(float(dur), int(seed), shift if len(shift) > 1 else shift[0][1])
        import zlib
        assert zlib.crc32(src_pcm.tobytes()) == int(g[name + '_pcm_crc'][0])
        assert zlib.crc32(dst_pcm.tobytes()) == int(g[name + '_pcm_crc'][1])
        dst = host_stream_from_pcm(dst_pcm, 12000, 1, 12000, stype)
        ref = oracle_stream_from_pcm(dst_pcm, 12000, 1, 12000, stype)
        assert np.array_equal(dst.data, ref.data)
        for off, n, center, window, d, t in g['{0}_{1}_calls'.format(name, stype)]:
            n = int(n)
            st, lo, span = dst._window(n, center, window)
            # oracle slicing (wav.py:178-184)
            from oracle.ref_matcher import clip
            st_ref = clip(center - window, -10, ref.duration_seconds)
            en_ref = clip(center + window, 0, ref.duration_seconds + 10)
            a = ref.sample_for_time(st_ref)
            b = ref.sample_for_time(en_ref) + n
            view = ref.data[:, a:b]
            assert st == st_ref and span == view.shape[1]
            assert lo == (view.__array_interface__['data'][0] - ref.data.__array_interface__['data'][0]) // ref.data.itemsize


def test_vectorised_planning_equals_scalar_path():
    """plan_queries (NumPy) against the scalar _window/_get_sample_for_time code, including windows
    clipped at both ends, negative times and the sub-sample rounding cases of int(rate * t)."""
    rng = np.random.default_rng(0)
    mk = lambda n: object.__new__(WavStream)
    src, dst = mk(0), mk(0)
    for s, count in ((src, 300.37), (dst, 299.2)):
        s.sample_rate, s.padding_size = 12000, 120000
        s.sample_count = int(np.ceil(count * 12000))
        s.data = np.zeros((1, 240000 + s.sample_count), np.uint8)
        s._handle = None
    starts = np.concatenate([rng.uniform(0, 295, 400), [0.0, 0.001, 299.0, 150.12345678]])
    starts = np.round(starts * 100) / 100
    ends = starts + rng.uniform(0.5, 4.0, len(starts))
    centers = starts + rng.uniform(-20, 20, len(starts))
    windows = rng.choice([1.5, 10.0, 60.0, 400.0], len(starts))
    toff, tlen, lag0, nlags, t0 = dst.plan_queries(src, starts, ends, centers, windows)
    for q in range(len(starts)):
        a, b = src._get_sample_for_time(starts[q]), src._get_sample_for_time(ends[q])
        lo, hi, _ = slice(a, b).indices(src.data.shape[1])
        st, l0, span = dst._window(hi - lo, centers[q], windows[q])
        assert (toff[q], tlen[q], lag0[q], nlags[q], t0[q]) == (lo, hi - lo, l0, span - (hi - lo) + 1, st)


def test_kernel_source_hash_ignores_comments_and_layout():
    """bench.kernel_source_hash stamps ncu captures: rewording a comment or re-indenting must not orphan a capture,
    changing code must."""
    import bench
    a = 'int f(int x) {\n    // add one\n    return x + 1;   /* done */\n}\n'
    b = 'int f(int x) {  // reworded\n\n  return x + 1;\n}'
    c = 'int f(int x) { return x + 2; }'
    s = 'const char* p = "// not a comment"; // a comment'
    assert bench._strip_comments(a) == bench._strip_comments(b) != bench._strip_comments(c)
    assert '"// not a comment"' in bench._strip_comments(s) and 'a comment' not in bench._strip_comments(s).replace('"// not a comment"', '')
    assert len(bench.kernel_source_hash()) == 16


def test_nvtx_ranges_are_optional_and_harmless(monkeypatch):
    """SUSHI_B200_NVTX=1 opens libnvToolsExt if there is one; without it, or with the variable unset, the ranges are
    no-ops.  Either way the bracketed code runs exactly once and exceptions pass through."""
    from sushi_b200 import _nvtx
    for flag in ('0', '1'):
        monkeypatch.setenv('SUSHI_B200_NVTX', flag)
        monkeypatch.setattr(_nvtx, '_tried', False)
        monkeypatch.setattr(_nvtx, '_lib', None)
        ran = []
        with _nvtx.nvtx_range('test range'):
            ran.append(1)
        assert ran == [1]
        with pytest.raises(ValueError):
            with _nvtx.nvtx_range('failing range'):
                raise ValueError('through')
        if flag == '0':
            assert not _nvtx.enabled()
