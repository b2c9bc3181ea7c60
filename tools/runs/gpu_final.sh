#!/bin/bash
# what the driver runs at round end, on one box
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
( time python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err ) 2>&1 | grep real; python -c "
import json; d=json.load(open('gpurun_out/final_bench_ref.json')); print(d['impl'], d['value'], d['cpu_baseline']['cores'])"
if [ $N -gt 1 ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_g$N.json 2> gpurun_out/final_bench_g$N.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench_g$N.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['spmv']['ms'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/final_bench_ref_g$N.json 2> gpurun_out/final_bench_ref_g$N.err; head -c 200 gpurun_out/final_bench_ref_g$N.json; echo
fi
