#!/bin/bash
# Short confirmation pass: GPU tests, the default bench line (with capture-stamped traffic), config 4.
T=${1:-r2h}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests -q -m gpu ) > $O/pytest_gpu_$T.txt 2>&1; tail -4 $O/pytest_gpu_$T.txt
timeout 600 python bench.py --no-load-leg > $O/bench_${T}_config3.json 2> $O/bench_${T}_config3.err; head -c 2600 $O/bench_${T}_config3.json; echo
timeout 300 python tools/config4.py 2>&1 | grep -v WARNING > $O/config4_$T.txt; tail -3 $O/config4_$T.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.txt 2>&1; tail -2 $O/smoke_$T.txt
