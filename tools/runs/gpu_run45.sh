#!/bin/bash
# coalescing microbenchmark (RED / store runs) + two-stage write-back: parity and timing
mkdir -p gpurun_out
timeout 120 tools/micro/build/redbench 2>&1 | grep -v "^bulk" > gpurun_out/r45_redbench_runs.txt; grep "runs" gpurun_out/r45_redbench_runs.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked_and_atomic" > gpurun_out/r45_pytest_a.log 2>&1; tail -5 gpurun_out/r45_pytest_a.log
for two in 1 0; do
  BFX_CHUNKS_TWO_STAGE=$two timeout 200 python bench.py --config p1 --no-cpu --no-e2e --spmv-reps 10 --steps 10 > gpurun_out/r45_bench_p1_two$two.json 2> gpurun_out/r45_bench_p1_two$two.err
  python -c "
import json; d=json.load(open('gpurun_out/r45_bench_p1_two$two.json')); print('p1 two$two', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['chunk_plan'])"
done
