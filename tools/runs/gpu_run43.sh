#!/bin/bash
# chunk-size sweep: P1 96/128/192 (256 and 384 measured in run 42), P2 64
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked_and_atomic" > gpurun_out/r43_pytest_a.log 2>&1; tail -3 gpurun_out/r43_pytest_a.log
for cb in 192 128 96; do
  BFX_CHUNKS_CB=$cb timeout 200 python bench.py --config p1 --no-cpu --no-e2e --spmv-reps 10 --steps 10 > gpurun_out/r43_bench_p1_cb$cb.json 2> gpurun_out/r43_bench_p1_cb$cb.err
  python -c "
import json; d=json.load(open('gpurun_out/r43_bench_p1_cb$cb.json')); print('p1 cb$cb', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['chunk_plan'])"
done
for cb in 64; do
  BFX_CHUNKS_CB=$cb timeout 200 python bench.py --config p2 --no-cpu --no-e2e --spmv-reps 10 --steps 10 > gpurun_out/r43_bench_p2_cb$cb.json 2> gpurun_out/r43_bench_p2_cb$cb.err
  python -c "
import json; d=json.load(open('gpurun_out/r43_bench_p2_cb$cb.json')); print('p2 cb$cb', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['chunk_plan'])"
done
