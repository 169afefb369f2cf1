#!/bin/bash
# GPU box (ONE GPU): the ncu passes behind profiles/r02_* (B200_PROFILING.md recipe).
# Numbers printed by programs running under ncu are not bench values.
set -u
OUT=gpurun_out/r02
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum"
# 1. launch list of the bench's steady state (shares of the three solver kernels)
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file $OUT/launches_bench_q1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-target > $OUT/bench_under_ncu.log 2>&1
# 2. --set full of the three Q1 solver kernels, steady state, PDL off (each kernel measured alone)
NBGPU_NO_PDL=1 ncu --set full --import-source on --cache-control none --clock-control none -k regex:krylov_ -s 600 -c 3 \
    -f -o $OUT/q1_solver_kernels python scripts/ab_pcg.py q1 > $OUT/ncu_q1.log 2>&1
# 3. K1 at Q16 and the SpMV at L64 (configs[3], configs[2]): traffic and throughput
ncu --metrics $M --cache-control none --clock-control none -k regex:krylov_spmv_stream -s 50 -c 2 --csv \
    --log-file $OUT/k1_q16.csv python scripts/ab_pcg.py q16 > $OUT/ncu_q16.log 2>&1
ncu --metrics $M --cache-control none --clock-control none -k regex:spmv_stream_kernel -s 10 -c 2 --csv \
    --log-file $OUT/spmv_l64.csv python scripts/spmv_sweep.py l8192 > $OUT/ncu_l64.log 2>&1
# 4. FEM-side kernels at Q1 (third repetition of each)
ncu --metrics $M --cache-control none --clock-control none -k regex:'assemble_|dirichlet_|strain_|stress_|gp_to_nodes|von_mises|vector_add' \
    -s 22 -c 11 --csv --log-file $OUT/fem_kernels.csv python scripts/fem_kernels.py > $OUT/ncu_fem.log 2>&1
ncu --set full --import-source on --cache-control none --clock-control none -k regex:assemble_node -s 2 -c 1 \
    -f -o $OUT/assemble_node_q1 python scripts/fem_kernels.py > $OUT/ncu_asm.log 2>&1
# 5. the row-partitioned kernel variants (PeerComm policy), one rank by itself
NBGPU_NO_PDL=1 ncu --metrics $M --cache-control none --clock-control none -k regex:'krylov_|halo_push' -s 300 -c 3 --csv \
    --log-file $OUT/dist_kernels_world1_q1.csv python scripts/dist_one.py q1 200 > $OUT/ncu_dist.log 2>&1
for f in $OUT/q1_solver_kernels $OUT/assemble_node_q1; do ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null; done
tail -2 $OUT/*.log | cut -c1-200
ls -la $OUT
