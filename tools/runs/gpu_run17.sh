#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv" > gpurun_out/pytest_spmv.log 2>&1; tail -3 gpurun_out/pytest_spmv.log
for cfg in p2; do
for v in 0 1 2; do
timeout 900 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --no-e2e --spmv-variant $v > gpurun_out/bench_${cfg}_sv$v.json 2> gpurun_out/bench_${cfg}_sv$v.err; tail -c 300 gpurun_out/bench_${cfg}_sv$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_sv$v.json'))
print('$cfg variant $v', 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']))
PY
done
done
