#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked or golden" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for cfg in p1 p2; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v25.json 2> gpurun_out/bench_${cfg}_v25.err; tail -c 300 gpurun_out/bench_${cfg}_v25.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v25.json'))
print(d.get("chunk_plan")); print("$cfg", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), d['setup_s'])
PY
done
