#!/usr/bin/env python3
"""ThreadSanitizer pass over the CPU emulation of the match kernels (tests/emu, test infrastructure).

    python tests/emu/run_tsan.py

Builds tests/emu/emu_driver.cpp with -fsanitize=thread and runs every kernel (one CTA per lag block / pair /
pair) with both screening loops, with and without the debug curve, on the cases of
tests/test_kernel_emulation.py.  Built with -DSB_EMU_THREADS: one OS thread stands for one CUDA thread (the default build runs the lanes of a warp as
fibers on one OS thread, which ThreadSanitizer cannot follow) and std::barrier for bar.sync, so a
shared-memory access that is not ordered by the kernel's own barriers / mbarrier waits shows up as a data race:
this checks the PLACEMENT of the barriers (re-use of the FFT buffer between the items of a pair,
re-staging of the sample windows, re-use of the reduction scratch), not the GPU memory model.  Exit code 0 and
"0 races" expected.  Re-executes itself under LD_PRELOAD=libtsan.so (the interpreter is not instrumented)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, 'tests', 'emu', '_build', 'libsb_emu_tsan.so')
LOG = os.path.join(ROOT, 'tests', 'emu', '_build', 'tsan.log')


def main():
    if os.environ.get('SB_EMU_TSAN_CHILD') != '1':
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(['g++', '-std=c++20', '-O1', '-g', '-fsanitize=thread', '-pthread', '-DSB_EMULATE', '-DSB_EMU_THREADS',
                               '-I', os.path.join(ROOT, 'tests', 'emu'), '-I', os.path.join(ROOT, 'sushi_b200', 'csrc'),
                               '-I', os.path.join(ROOT, 'include'), '-I', '/usr/local/cuda/include', '-shared', '-fPIC',
                               os.path.join(ROOT, 'tests', 'emu', 'emu_driver.cpp'), '-o', LIB])
        tsan = subprocess.check_output(['gcc', '-print-file-name=libtsan.so']).decode().strip()
        env = dict(os.environ, SB_EMU_TSAN_CHILD='1', LD_PRELOAD=tsan, TSAN_OPTIONS='halt_on_error=0 history_size=2 log_path=' + LOG)
        for f in os.listdir(os.path.dirname(LOG)):
            if f.startswith('tsan.log'):
                os.remove(os.path.join(os.path.dirname(LOG), f))
        rc = subprocess.call([sys.executable, os.path.abspath(__file__)], env=env)
        races = 0
        for f in os.listdir(os.path.dirname(LOG)):
            if f.startswith('tsan.log'):
                races += open(os.path.join(os.path.dirname(LOG), f)).read().count('WARNING: ThreadSanitizer')
        print('%d races reported (logs: %s.*), child exit code %d' % (races, LOG, rc))
        return 1 if races or rc else 0

    import numpy as np
    import tests.test_kernel_emulation as T
    lib = ctypes.CDLL(LIB)
    B = T.B
    n_img = 6 * B - 5000
    img = T.programme(n_img, 1)
    rng = np.random.default_rng(2)
    src = np.clip(np.roll(img, -700).astype(np.int32) + rng.integers(-5, 6, n_img), 0, 255).astype(np.uint8)
    case = T.Case(lib, img, src, [(30000, 20000, B + 300, 4 * B - 17000), (1000, 5000, 100, 20000),
                                  (8000, 40000, n_img - 40000 - 30000, 30001)], np.uint8)
    ref = None
    for kernel in (0, 1):
        for epi in (1, 3):
            for curves in (False, True):
                d, i, _ = case.run(kernel, epi, curves)
                ref = ref or (d, i)
                assert np.array_equal(ref[0], d) and np.array_equal(ref[1], i)
                print('kernel %d epilogue %d curves %d ok' % (kernel, epi, curves), flush=True)
    # the pair kernel is persistent: two CTAs walking all the pairs (state carried from one pair to the next)
    lib.emu_set_pair_grid(2)
    for epi in (1, 3):
        d, i, _ = case.run(1, epi, False)
        assert np.array_equal(ref[0], d) and np.array_equal(ref[1], i)
        print('persistent pair kernel, 2 CTAs, epilogue %d ok' % epi, flush=True)
    lib.emu_set_pair_grid(0)
    img32 = (T.programme(4 * B - 3000, 3).astype(np.float32) / 255.0).astype(np.float32)
    case32 = T.Case(lib, img32, np.roll(img32, -300).copy(), [(20000, 18000, 5, 2 * B + 5000)], np.float32)
    for kernel in (0, 1):
        d, i, _ = case32.run(kernel, 1, False)
        assert i[0] == 20295
        print('float32 kernel %d ok' % kernel, flush=True)
    return 0


if __name__ == '__main__':
    sys.exit(main())
