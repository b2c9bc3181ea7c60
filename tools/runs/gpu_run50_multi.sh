#!/bin/bash
# lean multi-GPU bench path: small check, then the C5 shard size (500^3 cells per GPU) on the GPUs of this box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" --no-cpu > gpurun_out/r50_bench_$name.json 2> gpurun_out/r50_bench_$name.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r50_bench_$name.json')); print('$name', d['n_gpus'], d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['spmv']['ms'], d['vector_assembly_ms'], d['sizes'], d['hbm'], d['setup_s'])
except Exception as e:
    print('$name failed', e); import subprocess; print(subprocess.run(['tail','-8','gpurun_out/r50_bench_$name.err'],capture_output=True,text=True).stdout)"
}
if [ "$1" != "big" ]; then
run g${N}_n128_lean --cells-per-edge 128 --lean --steps 10 --warmup 3 --spmv-reps 20
run g${N}_n128 --cells-per-edge 128 --no-e2e --steps 10 --warmup 3 --spmv-reps 20
fi
run g${N}_n500_lean --cells-per-edge 500 --lean --steps 5 --warmup 3 --spmv-reps 10
