#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q1 or elasticity or bcs" > gpurun_out/pytest_q1.log 2>&1; tail -5 gpurun_out/pytest_q1.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowgather" > gpurun_out/sanitizer_rowgather.log 2>&1; tail -4 gpurun_out/sanitizer_rowgather.log
for packed in 0 1; do
if [ $packed = 1 ]; then export BFX_ROWGATHER_PACKED=1; fi
timeout 900 python bench.py --config q1 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_q1_pk$packed.json 2> gpurun_out/bench_q1_pk$packed.err; tail -c 300 gpurun_out/bench_q1_pk$packed.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_q1_pk$packed.json'))
print("q1 packed=$packed", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), d['setup_s'])
PY
done
