mkdir -p gpurun_out/r2_41
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 400 --csv --log-file gpurun_out/r2_41/launches_p1.csv python bench.py --steps 3 --warmup 3 --legs none --no-cpu --no-e2e --no-parity > gpurun_out/r2_41/ncu_bench.log 2>&1
tail -2 gpurun_out/r2_41/ncu_bench.log | cut -c1-300
wc -l gpurun_out/r2_41/launches_p1.csv
