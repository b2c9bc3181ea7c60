#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_matrix_chunked" -s 3 -c 1 -o gpurun_out/prof_p1_128_chunked python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_chunked.log 2>&1; tail -3 gpurun_out/ncu_chunked.log
