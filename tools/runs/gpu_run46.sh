#!/bin/bash
# two-stage write-back, split lists: parity, timing, ncu capture with source counters
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked_and_atomic" > gpurun_out/r46_pytest_a.log 2>&1; tail -3 gpurun_out/r46_pytest_a.log
for two in 2; do
  BFX_CHUNKS_TWO_STAGE=$two timeout 200 python bench.py --config p1 --no-cpu --no-e2e --spmv-reps 10 --steps 10 > gpurun_out/r46_bench_p1_two$two.json 2> gpurun_out/r46_bench_p1_two$two.err
  python -c "
import json; d=json.load(open('gpurun_out/r46_bench_p1_two$two.json')); print('p1 two$two', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['chunk_plan'])"
done
BFX_CHUNKS_TWO_STAGE=2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_matrix_chunked" -s 3 -c 1 -o gpurun_out/prof_p1_256_two_stage_split python bench.py --config p1 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_46.log 2>&1; tail -1 gpurun_out/ncu_46.log | cut -c1-80
