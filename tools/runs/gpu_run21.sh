#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked and (p1_lex or tri)" > gpurun_out/sanitizer_chunked.log 2>&1; tail -4 gpurun_out/sanitizer_chunked.log
for cfg in p1 p2 q1; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v20.json 2> gpurun_out/bench_${cfg}_v20.err; tail -c 300 gpurun_out/bench_${cfg}_v20.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v20.json'))
print(d.get("chunk_plan")); print("$cfg", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'], d['setup_s'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_matrix_chunked" -s 3 -c 1 -o gpurun_out/prof_p1_256_chunked_v6 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_chunked.log 2>&1; tail -1 gpurun_out/ncu_chunked.log | cut -c1-100
