#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config q1 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_q1_192_v2.json 2> gpurun_out/bench_q1_192_v2.err; tail -c 500 gpurun_out/bench_q1_192_v2.err; cat gpurun_out/bench_q1_192_v2.json | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_stream" -c 2 -o gpurun_out/prof_spmv_p1_128 python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_blocked" -c 2 -o gpurun_out/prof_spmv_q1_96 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_elasticity_q1" -c 2 -o gpurun_out/prof_q1_96_v2 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -8
