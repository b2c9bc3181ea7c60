#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked and p2" > gpurun_out/pytest_p2.log 2>&1; tail -3 gpurun_out/pytest_p2.log
timeout 900 python bench.py --config p2 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_p2_v24.json 2> gpurun_out/bench_p2_v24.err; tail -c 300 gpurun_out/bench_p2_v24.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_p2_v24.json'))
print(d.get("chunk_plan")); print("p2", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), d['setup_s'])
PY
