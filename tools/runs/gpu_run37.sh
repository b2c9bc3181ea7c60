#!/bin/bash
# final ncu captures of the dominant kernels at the BASELINE sizes (traffic for roofline.traffic) + bench lines
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"k_matrix_chunked" -s 3 -c 1 -o gpurun_out/prof_p2_128_chunked python bench.py --config p2 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_1.log 2>&1; tail -1 gpurun_out/ncu_1.log | cut -c1-60
timeout 900 ncu --set full --clock-control none -k regex:"k_q1_rowgather" -s 3 -c 1 -o gpurun_out/prof_q1_192_rowgather python bench.py --config q1 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 1 > gpurun_out/ncu_2.log 2>&1; tail -1 gpurun_out/ncu_2.log | cut -c1-60
timeout 900 ncu --set full --clock-control none -k regex:"k_spmv_blocked_tma" -s 8 -c 1 -o gpurun_out/prof_spmv_q1_192_btma python bench.py --config q1 --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 > gpurun_out/ncu_3.log 2>&1; tail -1 gpurun_out/ncu_3.log | cut -c1-60
timeout 900 ncu --set full --clock-control none -k regex:"k_spmv_tma" -s 8 -c 1 -o gpurun_out/prof_spmv_p1_256_tma_v2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --spmv-reps 2 > gpurun_out/ncu_4.log 2>&1; tail -1 gpurun_out/ncu_4.log | cut -c1-60
for cfg in p1 p2 q1; do
timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${cfg}_final.json 2> gpurun_out/bench_${cfg}_final.err; tail -c 200 gpurun_out/bench_${cfg}_final.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${cfg}_final.json'))
print("$cfg", d["roofline"]["kernel"], 'step %.3f ms'%d['ms_per_step'], 'asm kernel %.3f ms frac %.3f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), 'spmv %.3f ms frac %.3f'%(d['spmv']['ms'], d['spmv']['frac']), 'vec %.3f lift %.3f'%(d['vector_assembly_ms'], d['apply_lifting_ms']))
PY
done
