#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_p1_n256_g${N}_v3.json 2> gpurun_out/bench_p1_n256_g${N}_v3.err; tail -c 300 gpurun_out/bench_p1_n256_g${N}_v3.err; cut -c1-330 gpurun_out/bench_p1_n256_g${N}_v3.json
