#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-3000
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | tail -3; cat gpurun_out/bench_reference.json | cut -c1-600
for cfg in p2 q1; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_v11.json 2> gpurun_out/bench_${cfg}_v11.err; tail -c 300 gpurun_out/bench_${cfg}_v11.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_v11.json'))
print('$cfg', d['roofline']['kernel'], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f'%d['vector_assembly_ms'])
PY
done
timeout 900 ncu --set full --clock-control none -k regex:"k_matrix_chunked" -s 3 -c 1 -o gpurun_out/prof_p1_256_chunked python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_p1_256.log 2>&1; tail -1 gpurun_out/ncu_p1_256.log | cut -c1-80
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_p1_256.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --spmv-reps 3 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | cut -c1-80
