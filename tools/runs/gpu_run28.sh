#!/bin/bash
mkdir -p gpurun_out
# launch list of the default bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_p1_256.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --spmv-reps 3 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | cut -c1-80; wc -l gpurun_out/launches_p1_256.csv
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err; cut -c1-600 gpurun_out/bench_default.json
