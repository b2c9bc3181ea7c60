#!/bin/bash
# round-end check on one GPU: all gpu tests, smoke, default bench
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r51_pytest_gpu.log 2>&1 ) 2>&1 | grep real; tail -4 gpurun_out/r51_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 300 python bench.py > gpurun_out/r51_bench.json 2> gpurun_out/r51_bench.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/r51_bench.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['e2e']['one_step_in_flight']['value'], d['cpu_baseline']['value'], d['clocks'], d['spmv']['frac'])"
