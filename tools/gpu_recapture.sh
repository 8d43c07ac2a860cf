#!/bin/bash
# After a change of the dominant kernel: bench lines, launch list and the full capture again (tests run elsewhere).
T=${1:-r2j}
O=gpurun_out
mkdir -p $O
python -c 'import bench; print(bench.kernel_source_hash())' > $O/src_hash_$T.txt
timeout 600 python bench.py > $O/bench_${T}_config3.json 2> $O/bench_${T}_config3.err; head -c 400 $O/bench_${T}_config3.json; echo
timeout 300 python bench.py --workload config2 --steps 20 --no-load-leg --cpu-budget 6 > $O/bench_${T}_config2.json 2>/dev/null; head -c 300 $O/bench_${T}_config2.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_launches_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 -o $O/pair_config3_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_full_$T.log 2>&1; tail -1 $O/ncu_full_$T.log
( time timeout 600 python tools/sweep.py --queries 1024 --out sweep_${T}_1gpu.json ) > $O/sweep_${T}_1gpu.txt 2>&1; tail -8 $O/sweep_${T}_1gpu.txt
