#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q1 or elasticity or bcs" > gpurun_out/pytest_q1.log 2>&1; tail -5 gpurun_out/pytest_q1.log
cfg=q1; st=auto
timeout 900 python bench.py --config $cfg --strategy $st --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_${st}_v10.json 2> gpurun_out/bench_${cfg}_${st}_v10.err; tail -c 400 gpurun_out/bench_${cfg}_${st}_v10.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_${st}_v10.json'))
print('$cfg $st', d['roofline']['kernel'], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'], d['setup_s'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_q1_rowgather" -s 3 -c 1 -o gpurun_out/prof_q1_96_rowgather_v3 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_rg.log 2>&1; tail -1 gpurun_out/ncu_rg.log | cut -c1-100
