#!/bin/bash
# weak scaling on N GPUs (N = visible GPUs): default bench config, device-timed, one process per GPU
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_g$N.log 2>&1; tail -3 gpurun_out/pytest_multi_g$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_p1_n256_g$N.json 2> gpurun_out/bench_p1_n256_g$N.err; tail -c 400 gpurun_out/bench_p1_n256_g$N.err; cat gpurun_out/bench_p1_n256_g$N.json | cut -c1-2200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_g$N.json 2> gpurun_out/bench_ref_g$N.err; tail -c 300 gpurun_out/bench_ref_g$N.err; cat gpurun_out/bench_ref_g$N.json | cut -c1-400
