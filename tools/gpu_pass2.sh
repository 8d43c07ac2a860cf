#!/bin/bash
# Second kind of GPU pass: A/B of prebuilt library variants (variants/lib_*.so), then tests + captures + the config-5
# sweep with the in-tree library.  Tag = $1.
T=${1:-r2b}
O=gpurun_out
mkdir -p $O
python -c 'import bench; print(bench.kernel_source_hash())' > $O/src_hash_$T.txt
bash tools/try_variants.sh > $O/variants_$T.txt 2>&1; cat $O/variants_$T.txt
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/pytest_gpu_$T.txt 2>&1; tail -4 $O/pytest_gpu_$T.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_launches_$T.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_pair -s 3 -c 1 -o $O/pair_config3_$T python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-load-leg > $O/ncu_full_$T.log 2>&1; tail -2 $O/ncu_full_$T.log
true
