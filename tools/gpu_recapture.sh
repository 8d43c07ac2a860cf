#!/bin/bash
# After a change of the dominant kernel: GPU tests, bench lines, launch list and the full capture again.
T=${1:-r2k}
O=gpurun_out
mkdir -p $O
python -c 'import bench; print(bench.kernel_source_hash())' > $O/src_hash_$T.txt
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/pytest_gpu_$T.txt 2>&1; grep -E "passed|failed" $O/pytest_gpu_$T.txt | tail -2
timeout 600 python bench.py > $O/bench_${T}_config3.json 2> $O/bench_${T}_config3.err; head -c 400 $O/bench_${T}_config3.json; echo
timeout 300 python bench.py --workload config2 --steps 20 --no-load-leg --cpu-budget 6 > $O/bench_${T}_config2.json 2>/dev/null; head -c 300 $O/bench_${T}_config2.json; echo
timeout 300 python bench.py --workload config2 --steps 10 --no-load-leg --no-cpu-baseline --sample-type float32 > $O/bench_${T}_config2_float32.json 2>/dev/null; head -c 200 $O/bench_${T}_config2_float32.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_launches_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 -o $O/pair_config3_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_full_$T.log 2>&1; tail -1 $O/ncu_full_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.txt 2>&1; tail -1 $O/smoke_$T.txt
