#!/bin/bash
# multi-GPU: parity tests + weak-scaling bench lines (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --n 128 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_p1_n128_g$N.json 2> gpurun_out/bench_p1_n128_g$N.err; tail -c 600 gpurun_out/bench_p1_n128_g$N.err; cat gpurun_out/bench_p1_n128_g$N.json | cut -c1-1500
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_p1_n256_g$N.json 2> gpurun_out/bench_p1_n256_g$N.err; tail -c 600 gpurun_out/bench_p1_n256_g$N.err; cat gpurun_out/bench_p1_n256_g$N.json | cut -c1-2500
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json | cut -c1-800
