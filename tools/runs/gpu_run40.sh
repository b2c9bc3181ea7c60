#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bcs or q1 or elasticity" > gpurun_out/pytest_q1.log 2>&1; tail -3 gpurun_out/pytest_q1.log
timeout 900 python bench.py --config q1 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_q1_final.json 2> gpurun_out/bench_q1_final.err; tail -c 200 gpurun_out/bench_q1_final.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q1_final.json'))
print("q1", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f lift %.3f'%(d['vector_assembly_ms'], d['apply_lifting_ms']))
PY
