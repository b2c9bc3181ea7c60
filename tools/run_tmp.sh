mkdir -p gpurun_out/r2_35
python -m pytest tests -m gpu -x -q > gpurun_out/r2_35/pytest.log 2>&1; tail -4 gpurun_out/r2_35/pytest.log
python bench.py > gpurun_out/r2_35/bench.json 2> gpurun_out/r2_35/bench.err; tail -c 400 gpurun_out/r2_35/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_35/bench.json').read().strip().splitlines()[-1])
def brief(l):
    print('  value %.3e ms/step %.3f kernel %.3f frac %.3f spmv %.3f(%.2f) vec %.3f lift %.3f parity %s plan %s hbm %s'%(l['value'],l['ms_per_step'],l['roofline']['kernel_ms'],l['roofline']['frac'],l['spmv']['ms'],l['spmv']['frac'],l['vector_assembly_ms'],l['apply_lifting_ms'],l['parity']['ok'], l['chunk_plan'], l['hbm']))
brief(d)
for k,v in d['configs'].items():
    print(k); brief(v)
PY
