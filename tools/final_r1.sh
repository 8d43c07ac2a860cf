#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun from the repo root): GPU tests, bench lines,
# ncu launch list + full capture of the dominant kernel on the bench workload, short config-5 sweep per engine.
# Everything lands in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
timeout 150 python bench.py > $O/bench_r1b.json 2> $O/bench_r1b.err; tail -c 400 $O/bench_r1b.json
for e in 1 3 5; do timeout 60 python bench.py --engine $e --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_r1b_engine$e.json 2>/dev/null; done
timeout 60 python bench.py --sample-type float32 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_r1b_float32.json 2>/dev/null
timeout 90 python bench.py --workload config3 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r1b_config3.json 2>/dev/null
timeout 60 python bench.py --workload config1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_r1b_config1.json 2>/dev/null
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 -o $O/pair_bench_r1b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
for e in 2; do timeout 90 python tools/sweep.py --engine $e --queries 128 --reps 2 --events 0.5,1,3,10,30 --windows 10,60,120 --out sweep_r1b_engine$e.json > $O/sweep_engine$e.txt 2>&1; tail -5 $O/sweep_engine$e.txt; done
timeout 120 python tests/parity_report.py > $O/parity_r1b.txt 2>&1; tail -4 $O/parity_r1b.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
