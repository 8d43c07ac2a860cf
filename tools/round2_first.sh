#!/bin/bash
# First GPU pass of round 2 (run under gpurun from the repo root, one B200): does the default path still pass,
# do the two opt-in variants written blind at the end of round 1 (sb_set_epilogue(2), engine 6) give bit-identical
# results, and what do they cost.  Everything lands in gpurun_out/.
# Parts of the second kernel body can be compiled out to attribute a difference (rebuild HERE, the .so travels):
#   touch sushi_b200/csrc/sb_fused2.cu && make -C sushi_b200/csrc EXTRA="-DSB_V2_MIDBAR=0"            (or =1 default)
#   EXTRA="-DSB_V2_MIDBAR_SPLIT=8"   EXTRA="-DSB_V2_SPECIAL_PREFETCH=0"
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
SB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu > $O/pytest_experimental.txt 2>&1; tail -15 $O/pytest_experimental.txt
B="--steps 10 --warmup 3 --no-cpu-baseline"
for v in "engine4_epi1 --engine 4 --epilogue 1" "engine4_epi2 --engine 4 --epilogue 2" "engine5_epi2 --engine 5 --epilogue 2" \
         "engine6_epi1 --engine 6 --epilogue 1" "engine6_epi2 --engine 6 --epilogue 2" \
         "engine4_epi1_bfp --engine 4 --epilogue 1 --spectra 1" "engine6_epi2_bfp --engine 6 --epilogue 2 --spectra 1" "engine5_epi2_bfp --engine 5 --epilogue 2 --spectra 1"; do
    set -- $v; name=$1; shift
    timeout 90 python bench.py $B "$@" > $O/bench_r2_$name.json 2> $O/bench_r2_$name.err
    python - "$O/bench_r2_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
    print(sys.argv[2], d['value'], 'events/s', d['roofline']['kernel_ms_per_step'].get('match_fused'), 'ms match kernel; shift mismatches', d['shift_check']['mismatches'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
done
timeout 60 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline --engine 6 --epilogue 2 > $O/bench_r2_config3_engine6_epi2.json 2>/dev/null
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2_first.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --engine 6 --epilogue 2 > $O/ncu_launches.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k_match_triple -s 3 -c 1 -o $O/triple_bench_r2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --engine 6 --epilogue 2 > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
SB_PARITY_EXPERIMENTAL=1 timeout 200 python tests/parity_report.py > $O/parity_r2_first.txt 2>&1; grep -c same $O/parity_r2_first.txt; grep DIFF $O/parity_r2_first.txt | head
timeout 120 python tools/sweep.py --engine 6 --epilogue 2 --queries 128 --reps 2 --events 0.5,1,3,10 --windows 10,60,120 --out sweep_r2_engine6_epi2.json > $O/sweep_engine6.txt 2>&1; tail -5 $O/sweep_engine6.txt
timeout 60 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_config3_default.json 2>/dev/null
timeout 60 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline --engine 4 --epilogue 2 > $O/bench_r2_config3_engine4_epi2.json 2>/dev/null
head -c 600 $O/bench_r2_config3_*.json
