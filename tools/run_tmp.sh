mkdir -p gpurun_out/r2_29
B="python bench.py --config p2 --steps 10 --warmup 3 --no-cpu --no-e2e --legs none"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_29/p2_$name.json 2> gpurun_out/r2_29/p2_$name.err; python - $name <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2_29/p2_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1],'kernel %.3f frac %.3f step %.3f parity %s plan_bytes %s'%(d['roofline']['kernel_ms'],d['roofline']['frac'],d['ms_per_step'],d['parity'] and d['parity']['ok'], d['chunk_plan']['plan_bytes']))
except Exception as e:
    print(sys.argv[1],'FAILED',e); print(open('gpurun_out/r2_29/p2_%s.err'%sys.argv[1]).read()[-1500:])
PY
}
run def A=1
run cb96 BFX_CHUNKS_CB=96
run cb64 BFX_CHUNKS_CB=64
