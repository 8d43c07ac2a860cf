"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference algorithm for the audio template-matching path
(tp7/Sushi wav.py + the part of sushi.py that drives it).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the reported CPU baseline -- never as part
of the product path (sushi_b200/ does not import it and has no CPU fallback).

Where the arithmetic lives: the reference calls a third-party dependency that is not
under /root/reference -- OpenCV (cv2.matchTemplate wav.py:185, cv2.resize wav.py:133)
and NumPy; the reference pins neither ("OpenCV 2.4.x or newer", "NumPy 1.8 or newer",
README.md:30-31; requirements.txt lists only numpy, mock).  The oracle therefore calls
THIS image's cv2 4.13.0 / numpy 2.3.5 at the same call sites.

Parity pin: the reference's own tests hold no golden vectors for this path
(hot-path coverage zero).  The oracle is pinned instead against outputs of
the reference itself run in this container: oracle/gen_golden.py imports
/root/reference/wav.py under Python 3 and runs its unmodified get_substream /
find_substream / WavStream.__init__ (with a bytes-compat shim for the RIFF reader) and
an in-memory py2->py3 transform of sushi.py's calculate_shifts; the results are frozen in
tests/golden/*.npz and tests/test_oracle_golden.py checks this restatement against them.
"""
