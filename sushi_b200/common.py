"""Small helpers shared by the host side (mirror of the pieces of the reference's
common.py that the hot path touches: SushiError common.py:4, clip common.py:41-42,
format_time common.py:32-38 -- used only in log lines)."""


class SushiError(Exception):
    """User-facing failure; the reference's CLI prints it and exits 2 (sushi.py:841-843)."""


def clip(value, minimum, maximum):
    # reference common.py:41-42 -- note the order: min() first, then max()
    return max(min(value, maximum), minimum)


def py2_round(x):
    """round() as Python 2 does it (half away from zero): the reference is py2 code
    (wav.py:127, common.py:33, subs.py:116)."""
    import math
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def format_time(seconds):
    cs = py2_round(seconds * 100)
    return '{0}:{1:02d}:{2:02d}.{3:02d}'.format(
        int(cs // 360000), int((cs // 6000) % 60), int((cs // 100) % 60), int(cs % 100))
