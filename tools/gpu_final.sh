#!/bin/bash
# Final 1-GPU pass of a round (run under gpurun from the repo root): full GPU tests, the bench lines, launch list,
# full captures of the dominant kernel and of the loader / scan kernels, config-5 sweep, parity report, config 4.
# Everything lands in gpurun_out/ under the tag given as $1.
T=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$T.txt
python -c 'import bench; print(bench.kernel_source_hash())' > $O/src_hash_$T.txt
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/pytest_gpu_$T.txt 2>&1; tail -4 $O/pytest_gpu_$T.txt
timeout 600 python bench.py > $O/bench_${T}_config3.json 2> $O/bench_${T}_config3.err; tail -2 $O/bench_${T}_config3.err; head -c 600 $O/bench_${T}_config3.json; echo
timeout 300 python bench.py --workload config2 --steps 20 --no-load-leg --cpu-budget 6 > $O/bench_${T}_config2.json 2> $O/bench_${T}_config2.err; head -c 300 $O/bench_${T}_config2.json; echo
timeout 300 python bench.py --workload config2 --steps 10 --no-load-leg --no-cpu-baseline --sample-type float32 > $O/bench_${T}_config2_float32.json 2>/dev/null; head -c 200 $O/bench_${T}_config2_float32.json; echo
timeout 300 python bench.py --workload config1 --steps 20 --no-load-leg --cpu-budget 4 > $O/bench_${T}_config1.json 2>/dev/null; head -c 200 $O/bench_${T}_config1.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_launches_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 -o $O/pair_config3_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_full_$T.log 2>&1; tail -1 $O/ncu_full_$T.log
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_decode_resample_pad|k_select_fine|k_normalise|k_tile_totals_u8|k_tile_scan_u8' -c 5 -o $O/loader_$T python bench.py --load-only --no-cpu-baseline > $O/ncu_loader_$T.log 2>&1; tail -1 $O/ncu_loader_$T.log
( time timeout 600 python tools/sweep.py --queries 1024 --out sweep_${T}_1gpu.json ) > $O/sweep_${T}_1gpu.txt 2>&1; tail -10 $O/sweep_${T}_1gpu.txt
timeout 400 python tests/parity_report.py > $O/parity_$T.txt 2>&1; tail -3 $O/parity_$T.txt
timeout 300 python tools/config4.py > $O/config4_$T.txt 2>&1; tail -4 $O/config4_$T.txt
