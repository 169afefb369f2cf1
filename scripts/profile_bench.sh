#!/bin/bash
# GPU box: the two ncu passes the profiles/ records come from (B200_PROFILING.md recipe), over the bench command.
#   1. launch list (one metric, no clock control) of 300 steady-state launches
#   2. --set full capture of the three solver kernels (PDL off so that every kernel is measured alone)
# Numbers printed by bench.py under ncu are not bench values.
set -u
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.log 2>&1
NBGPU_NO_PDL=1 ncu --set full --import-source on --cache-control none --clock-control none \
    -k regex:krylov_ -s 600 -c 3 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
tail -3 $OUT/ncu_full_$TAG.log | cut -c1-200
head -5 $OUT/launches_$TAG.csv | cut -c1-200
