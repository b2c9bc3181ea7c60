#!/bin/bash
# chunk-aggregated assembly: parity, memcheck of the plan + kernel, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked" > gpurun_out/pytest_chunked.log 2>&1; tail -15 gpurun_out/pytest_chunked.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked and p1_lex" > gpurun_out/sanitizer_chunked.log 2>&1; tail -8 gpurun_out/sanitizer_chunked.log
for cfg in p1 p2; do
for st in auto atomic; do
timeout 900 python bench.py --config $cfg --strategy $st --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_${st}_v5.json 2> gpurun_out/bench_${cfg}_${st}_v5.err; tail -c 400 gpurun_out/bench_${cfg}_${st}_v5.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_${st}_v5.json'))
print('$cfg $st', d['roofline']['kernel'], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'], d['setup_s'])
PY
done
done
