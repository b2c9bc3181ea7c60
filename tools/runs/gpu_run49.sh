#!/bin/bash
# the C5 shard (500^3 cells per GPU) with the default (chunked) strategy
mkdir -p gpurun_out
name=p1_n500_chunked
timeout 500 python bench.py --config p1 --n 500 --no-cpu --no-e2e --spmv-reps 10 --steps 5 --warmup 3 > gpurun_out/r49_bench_$name.json 2> gpurun_out/r49_bench_$name.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/r49_bench_$name.json')); print('$name', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['spmv']['ms'], d['spmv']['frac'], d['sizes'], d['hbm'], d['setup_s'], d['chunk_plan'])
except Exception as e:
    print('$name failed', e); import subprocess; print(subprocess.run(['tail','-5','gpurun_out/r49_bench_$name.err'],capture_output=True,text=True).stdout)"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
