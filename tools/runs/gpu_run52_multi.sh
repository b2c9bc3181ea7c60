#!/bin/bash
# 2-GPU check of what the driver's scaling run executes: multi-GPU tests + default bench under torchrun
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r52_bench_g$N.json 2> gpurun_out/r52_bench_g$N.err; python -c "
import json; d=json.load(open('gpurun_out/r52_bench_g$N.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d['spmv']['ms'])" || tail -5 gpurun_out/r52_bench_g$N.err
timeout 120 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r52_pytest_multi.log 2>&1; tail -3 gpurun_out/r52_pytest_multi.log
