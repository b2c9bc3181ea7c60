mkdir -p gpurun_out/r2_42
python -m pytest tests -m gpu -x -q -k "q1 or Q1 or elasticity" > gpurun_out/r2_42/pytest_q1.log 2>&1; tail -3 gpurun_out/r2_42/pytest_q1.log
python bench.py --config q1 --steps 10 --warmup 3 --no-cpu --no-e2e --legs none > gpurun_out/r2_42/q1.json 2> gpurun_out/r2_42/q1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_42/q1.json').read().strip().splitlines()[-1])
print('q1 kernel %.3f frac %.3f step %.3f parity %s'%(d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['parity']['ok']))
PY
