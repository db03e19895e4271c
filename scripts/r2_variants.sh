# A/B timing of the actor kernel's build variants (TACTOR_VARIANT), small bridge 4096 and large bridge 8192
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2 3 4}; do
  TACTOR_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.1 > gpurun_out/r2_bench_v$v.json 2> gpurun_out/r2_bench_v$v.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_v$v.json').readlines()[-1]);print('variant $v small', d['value'], d['stages']['actor_ms'])
except Exception as e: print('variant $v failed', e, open('gpurun_out/r2_bench_v$v.err').read()[-1500:])"
  TACTOR_VARIANT=$v timeout 300 python bench.py --family large_bridge --batch 8192 --steps 20 --warmup 5 --cpu-seconds 0.1 > gpurun_out/r2_bench_large_v$v.json 2>/dev/null
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_large_v$v.json').readlines()[-1]);print('variant $v large', d['value'], d['stages']['actor_ms'])
except Exception as e: print('large variant $v failed', e)"
done
