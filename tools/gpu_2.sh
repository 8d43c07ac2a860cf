#!/bin/bash
# 2-GPU check (gpurun --gpus 2): parity of the sharded path and the default bench at 2 ranks.
T=${1:-r2m}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/check_sharded_gpu.py > $O/sharded2_$T.txt 2>&1; tail -2 $O/sharded2_$T.txt
timeout 300 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_${T}_2gpu_config3.json 2> $O/bench_${T}_2gpu_config3.err; head -c 500 $O/bench_${T}_2gpu_config3.json; echo
