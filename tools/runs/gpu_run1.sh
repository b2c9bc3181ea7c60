#!/bin/bash
# first GPU run: parity tests, bench, launch list, full ncu capture of the top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo skip-tests

timeout 300 python bench.py --n 64 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_p1_64.json 2> gpurun_out/bench_p1_64.err; tail -c 600 gpurun_out/bench_p1_64.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_p1_256.json 2> gpurun_out/bench_p1_256.err; tail -c 1500 gpurun_out/bench_p1_256.err
cat gpurun_out/bench_p1_256.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_p1_128.csv python bench.py --n 128 --steps 2 --warmup 3 --no-cpu --no-e2e --spmv-reps 3 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_matrix_cells|k_spmv_stream|k_vector_cells" -c 6 -o gpurun_out/prof_p1_128 python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
