#!/bin/bash
# 5 resident CTAs per SM (48 registers) for the P1 chunked kernel: parity + timing
mkdir -p gpurun_out
BFX_CHUNK_OCC=5 timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -k "poisson_p1 or (chunked_and_atomic and p1)" > gpurun_out/r53_pytest.log 2>&1; tail -2 gpurun_out/r53_pytest.log
BFX_CHUNK_OCC=5 timeout 100 python bench.py --config p1 --no-cpu --no-e2e --spmv-reps 5 --steps 10 > gpurun_out/r53_bench_p1_occ5.json 2> gpurun_out/r53_bench_p1_occ5.err
python -c "
import json; d=json.load(open('gpurun_out/r53_bench_p1_occ5.json')); print('p1 occ5', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
