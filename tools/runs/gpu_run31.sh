#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vector or matrix_free or golden or config1 or bcs" > gpurun_out/pytest_vec.log 2>&1; tail -12 gpurun_out/pytest_vec.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vector_assembly_grouped" > gpurun_out/sanitizer_vec.log 2>&1; tail -4 gpurun_out/sanitizer_vec.log
for gv in 1 0; do
BFX_GROUPED_VECTORS=$gv timeout 900 python bench.py --config p1 --steps 3 --warmup 3 --no-cpu --no-e2e --spmv-reps 5 > gpurun_out/bench_p1_gv$gv.json 2> gpurun_out/bench_p1_gv$gv.err; tail -c 300 gpurun_out/bench_p1_gv$gv.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_p1_gv$gv.json'))
print("p1 grouped=$gv", 'vec %.3f ms'%d['vector_assembly_ms'])
PY
done
