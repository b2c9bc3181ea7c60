#!/bin/bash
# new tests (full-size properties, two steps in flight, alternative chunk size) + A/B of the chunk size
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked_and_atomic or two_in_flight or host_buffer" > gpurun_out/r42_pytest_a.log 2>&1; tail -3 gpurun_out/r42_pytest_a.log
( time timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/r42_pytest_full.log 2>&1 ) 2>&1 | grep real; tail -15 gpurun_out/r42_pytest_full.log
for alt in 0 1; do
  BFX_CHUNKS_ALT_CB=$alt timeout 200 python bench.py --config p1 --no-cpu --spmv-reps 10 $( [ $alt = 1 ] && echo --no-e2e ) > gpurun_out/r42_bench_p1_alt$alt.json 2> gpurun_out/r42_bench_p1_alt$alt.err
  python -c "
import json; d=json.load(open('gpurun_out/r42_bench_p1_alt$alt.json')); print('p1 alt$alt', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'], d['chunk_plan'])"
  BFX_CHUNKS_ALT_CB=$alt timeout 200 python bench.py --config p2 --no-cpu --no-e2e --spmv-reps 10 > gpurun_out/r42_bench_p2_alt$alt.json 2> gpurun_out/r42_bench_p2_alt$alt.err
  python -c "
import json; d=json.load(open('gpurun_out/r42_bench_p2_alt$alt.json')); print('p2 alt$alt', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['chunk_plan'])"
done
