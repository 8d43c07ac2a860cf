#!/bin/bash
# compute-sanitizer over a small workload of the default path (and the float32 path), plus full captures of the two
# small kernels next to the dominant one.
T=${1:-r2n}
O=gpurun_out
mkdir -p $O
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/ncu_case.py --events 60 --duration 120 --window 20 --batches 1 > $O/sanitizer_memcheck_u8_$T.txt 2>&1; echo "memcheck u8 rc=$?"; tail -3 $O/sanitizer_memcheck_u8_$T.txt
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/ncu_case.py --events 40 --duration 120 --window 20 --batches 1 --sample-type float32 > $O/sanitizer_memcheck_f32_$T.txt 2>&1; echo "memcheck f32 rc=$?"; tail -3 $O/sanitizer_memcheck_f32_$T.txt
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/ncu_case.py --events 12 --duration 60 --window 10 --batches 1 > $O/sanitizer_racecheck_u8_$T.txt 2>&1; echo "racecheck u8 rc=$?"; tail -3 $O/sanitizer_racecheck_u8_$T.txt
