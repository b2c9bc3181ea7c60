#!/bin/bash
mkdir -p gpurun_out
for sym in 1 0; do
for cfg in p1 p2; do
BFX_CHUNKS_SYMMETRIC=$sym timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_sym$sym.json 2> gpurun_out/bench_${cfg}_sym$sym.err; tail -c 300 gpurun_out/bench_${cfg}_sym$sym.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_sym$sym.json'))
print(d.get("chunk_plan")); print("$cfg sym=$sym", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']))
PY
done
done
