#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv" > gpurun_out/pytest_spmv.log 2>&1; tail -3 gpurun_out/pytest_spmv.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv_assembled and q1" > gpurun_out/sanitizer_spmv.log 2>&1; tail -4 gpurun_out/sanitizer_spmv.log
for v in 0 1; do
timeout 900 python bench.py --config q1 --steps 3 --warmup 3 --no-cpu --no-e2e --spmv-variant $v > gpurun_out/bench_q1_sv$v.json 2> gpurun_out/bench_q1_sv$v.err; tail -c 300 gpurun_out/bench_q1_sv$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q1_sv$v.json'))
print('q1 variant $v', 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']))
PY
done
