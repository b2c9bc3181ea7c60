#!/bin/bash
# last sanity of the default path after the OCC template parameter: P1 parity subset + quick bench
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "poisson_p1 or (chunked_and_atomic and p1_lex) or transpose" > gpurun_out/r54_pytest.log 2>&1; tail -2 gpurun_out/r54_pytest.log
timeout 60 python bench.py --config p1 --no-cpu --no-e2e --spmv-reps 5 --steps 10 > gpurun_out/r54_bench_p1.json 2> gpurun_out/r54_bench_p1.err
python -c "
import json; d=json.load(open('gpurun_out/r54_bench_p1.json')); print('p1', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
