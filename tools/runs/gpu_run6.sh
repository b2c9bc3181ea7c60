#!/bin/bash
# validation of the unique-pair Q1 kernel + SpMV variants, microbenchmarks for the scatter-add choice
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/redbench tools/micro/redbench.cu && timeout 120 /tmp/redbench > gpurun_out/redbench.txt 2>&1; cat gpurun_out/redbench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for cfg in p1 q1; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v4.json 2> gpurun_out/bench_${cfg}_v4.err; tail -c 300 gpurun_out/bench_${cfg}_v4.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v4.json'))
print('$cfg', 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'])
PY
done
