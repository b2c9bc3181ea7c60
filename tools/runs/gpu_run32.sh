#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vector_grouped" -s 2 -c 1 -o gpurun_out/prof_vec_grouped python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_v.log 2>&1; tail -1 gpurun_out/ncu_v.log | cut -c1-80
