#!/bin/bash
# shard sizes beyond C2 on one GPU: 400^3 (chunked) and the C5 shard 500^3 (RED kernel: the chunk plan does not fit 180 GB)
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 400 python bench.py "$@" --no-cpu --no-e2e --spmv-reps 10 --steps 5 --warmup 3 > gpurun_out/r48_bench_$name.json 2> gpurun_out/r48_bench_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r48_bench_$name.json')); print('$name', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['spmv']['ms'], d['spmv']['frac'], d['sizes'], d['hbm'], d['setup_s'])
except Exception as e:
    print('$name failed', e); import subprocess; print(subprocess.run(['tail','-3','gpurun_out/r48_bench_$name.err'],capture_output=True,text=True).stdout)"
}
run p1_default_check --config p1
run p1_n400 --config p1 --n 400
run p1_n500_atomic --config p1 --n 500 --strategy atomic
