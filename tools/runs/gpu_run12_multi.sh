#!/bin/bash
# multi-GPU: parity tests (bounded) + weak-scaling bench lines (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi.log 2>&1; tail -30 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_p1_n256_g$N.json 2> gpurun_out/bench_p1_n256_g$N.err; tail -c 600 gpurun_out/bench_p1_n256_g$N.err; cat gpurun_out/bench_p1_n256_g$N.json | cut -c1-1800
