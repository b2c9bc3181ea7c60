#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for cfg in p1 p2 q1; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v3.json 2> gpurun_out/bench_${cfg}_v3.err; tail -c 300 gpurun_out/bench_${cfg}_v3.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v3.json'))
print('$cfg', 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_stream" -c 2 -o gpurun_out/prof_spmv_p1_128_v3 python bench.py --n 128 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_blocked" -c 2 -o gpurun_out/prof_spmv_q1_96_v3 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_elasticity_q1" -c 2 -o gpurun_out/prof_q1_96_v3 python bench.py --config q1 --n 96 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_c.log 2>&1
