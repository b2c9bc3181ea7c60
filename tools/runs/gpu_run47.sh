#!/bin/bash
# L2 prefetch of the next wave's plan streams / coordinates: parity (prefetch on) and timing
mkdir -p gpurun_out
BFX_CHUNK_PREFETCH=2 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chunked_and_atomic or poisson_p1 or poisson_p2" > gpurun_out/r47_pytest_a.log 2>&1; tail -3 gpurun_out/r47_pytest_a.log
run() { # name, config, env...
  name=$1; cfg=$2; shift 2
  env "$@" timeout 200 python bench.py --config $cfg --no-cpu --no-e2e --spmv-reps 10 --steps 10 > gpurun_out/r47_bench_$name.json 2> gpurun_out/r47_bench_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r47_bench_$name.json')); print('$name', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
}
run p1_pf1 p1 BFX_CHUNK_PREFETCH=1
run p1_pf2 p1 BFX_CHUNK_PREFETCH=2
run p1_pf2_d1184 p1 BFX_CHUNK_PREFETCH=2 BFX_CHUNK_PREFETCH_DIST=1184
run p1_pf1_d296 p1 BFX_CHUNK_PREFETCH=1 BFX_CHUNK_PREFETCH_DIST=296
run p2_pf0 p2 BFX_CHUNK_PREFETCH=0
run p2_pf1 p2 BFX_CHUNK_PREFETCH=1
run p2_pf2 p2 BFX_CHUNK_PREFETCH=2
