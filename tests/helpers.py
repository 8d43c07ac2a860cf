"""Shared helpers for the parity tests: oracle-side stream construction from PCM."""
import numpy as np

from oracle import ref_loader, ref_matcher


def oracle_stream_from_pcm(pcm, framerate, channels, sample_rate, sample_type):
    """PCM -> RefStream through the oracle's restatement of the reference loader."""
    raw = np.ascontiguousarray(pcm, '<i2').reshape(-1).view(np.uint8)
    pos = [0]
    frame_size = 2 * channels

    def read_raw(nframes):
        a = pos[0]
        b = min(a + nframes * frame_size, raw.size)
        pos[0] = b
        return raw[a:b].tobytes()
    data, sample_count, padding = ref_loader.load_stream(read_raw, (raw.size // frame_size), framerate, 2, channels,
                                                         sample_rate, sample_type)
    return ref_matcher.RefStream(data, sample_rate, padding, sample_count)
