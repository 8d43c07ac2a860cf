#!/bin/bash
# A/B of prebuilt library variants (variants/lib_*.so) on one box, then the GPU tests with the in-tree library.  Tag = $1.
T=${1:-r2b}
O=gpurun_out
mkdir -p $O
python -c 'import bench; print(bench.kernel_source_hash())' > $O/src_hash_$T.txt
bash tools/try_variants.sh > $O/variants_$T.txt 2>&1; cat $O/variants_$T.txt
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/pytest_gpu_$T.txt 2>&1; tail -4 $O/pytest_gpu_$T.txt
timeout 200 python tools/config4.py 2>&1 | grep -v WARNING > $O/config4_$T.txt; tail -3 $O/config4_$T.txt
