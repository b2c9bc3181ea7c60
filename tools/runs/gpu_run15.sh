#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_rows|k_spmv_stream" -s 8 -c 2 -o gpurun_out/prof_spmv_p1_256_rows python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-80
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv" -s 8 -c 2 -o gpurun_out/prof_spmv_p2_128 python bench.py --config p2 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-80
