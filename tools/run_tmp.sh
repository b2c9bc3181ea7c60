mkdir -p gpurun_out/r2_45
python -m pytest tests -m gpu -x -q > gpurun_out/r2_45/pytest.log 2>&1; tail -3 gpurun_out/r2_45/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --legs none --no-parity > gpurun_out/r2_45/p1.json 2> gpurun_out/r2_45/p1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_45/p1.json').read().strip().splitlines()[-1])
print('vec %.3f lift %.3f step %.3f'%(d['vector_assembly_ms'], d['apply_lifting_ms'], d['ms_per_step']), d['hbm'])
PY
