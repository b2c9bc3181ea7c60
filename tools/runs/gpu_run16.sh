#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv" > gpurun_out/pytest_spmv.log 2>&1; tail -5 gpurun_out/pytest_spmv.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv_assembled and p1" > gpurun_out/sanitizer_spmv.log 2>&1; tail -4 gpurun_out/sanitizer_spmv.log
for cfg in p1 p2; do
timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v12.json 2> gpurun_out/bench_${cfg}_v12.err; tail -c 300 gpurun_out/bench_${cfg}_v12.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v12.json'))
print('$cfg', d['roofline']['kernel'], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_tma" -s 2 -c 1 -o gpurun_out/prof_spmv_p1_256_tma python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log | cut -c1-80
