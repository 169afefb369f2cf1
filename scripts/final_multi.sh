#!/bin/bash
# GPU box, N GPUs: the bench lines of profiles/r02_bench_n<N>*.json.   scripts/final_multi.sh N
N=$1
mkdir -p gpurun_out/fin
T="timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
$T 29511 bench.py --gpus $N > gpurun_out/fin/b$N.log 2>&1
NBGPU_PCG_MODE=fused $T 29513 bench.py --gpus $N --no-target > gpurun_out/fin/b${N}_fused.log 2>&1
tail -c 300 gpurun_out/fin/b$N.log
