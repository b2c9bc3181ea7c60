#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config p2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_p2_128.json 2> gpurun_out/bench_p2_128.err; tail -c 800 gpurun_out/bench_p2_128.err
timeout 900 python bench.py --config q1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_q1_192.json 2> gpurun_out/bench_q1_192.err; tail -c 800 gpurun_out/bench_q1_192.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_stream|k_vector_cells" -c 4 -o gpurun_out/prof_p1_128_spmv python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_elasticity_q1|k_spmv_blocked" -c 4 -o gpurun_out/prof_q1_96 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_full3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_matrix_cells" -c 2 -o gpurun_out/prof_p2_64 python bench.py --config p2 --n 64 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out
