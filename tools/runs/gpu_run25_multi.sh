#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_g$N.log 2>&1; tail -3 gpurun_out/pytest_multi_g$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_p1_n256_g${N}_v2.json 2> gpurun_out/bench_p1_n256_g${N}_v2.err; tail -c 400 gpurun_out/bench_p1_n256_g${N}_v2.err; cat gpurun_out/bench_p1_n256_g${N}_v2.json | cut -c1-400
