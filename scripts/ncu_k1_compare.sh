#!/bin/bash
# ncu --set full of ONE steady-state K1 launch on the single-GPU path and on the world=1 row-partitioned path
SZ=${1:-q4}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:krylov_spmv_stream -s 40 -c 1 -f -o gpurun_out/k1_single_$SZ \
    python scripts/ab_pcg.py $SZ > gpurun_out/ncu_single.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:krylov_spmv_stream -s 40 -c 1 -f -o gpurun_out/k1_dist_$SZ \
    python scripts/dist_one.py $SZ 100 > gpurun_out/ncu_dist.log 2>&1
tail -2 gpurun_out/ncu_single.log gpurun_out/ncu_dist.log
ls -la gpurun_out/*.ncu-rep
