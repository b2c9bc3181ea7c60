#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_blocked_tma" -s 2 -c 1 -o gpurun_out/prof_spmv_q1_96_btma python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 --spmv-variant 1 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-80
