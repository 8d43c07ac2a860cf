#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): parity of the sharded path, the config-5 sweep and the default bench at 8 ranks.
T=${1:-r2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L > $O/smi8_$T.txt
timeout 300 $TR --master-port 29511 tools/check_sharded_gpu.py > $O/sharded8_$T.txt 2>&1; tail -3 $O/sharded8_$T.txt
timeout 400 $TR --master-port 29512 tools/sweep.py --queries 1024 --out sweep_${T}_8gpu.json > $O/sweep_${T}_8gpu.txt 2>&1; tail -9 $O/sweep_${T}_8gpu.txt
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_${T}_8gpu_config3.json 2> $O/bench_${T}_8gpu_config3.err; head -c 700 $O/bench_${T}_8gpu_config3.json; echo
timeout 300 $TR --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 --workload config2 --scaling weak > $O/bench_${T}_8gpu_config2_weak.json 2>/dev/null; head -c 300 $O/bench_${T}_8gpu_config2_weak.json; echo
