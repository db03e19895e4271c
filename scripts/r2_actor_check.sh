# first GPU check of the round-2 actor kernel (run under gpurun, 1 GPU); outputs in gpurun_out/
mkdir -p gpurun_out
python -c "from mop_truss_marl_b200 import actor; actor.selftest_tmem_layout(0); print('tmem layout ok')" > gpurun_out/r2_selftest.log 2>&1
tail -2 gpurun_out/r2_selftest.log
timeout 1200 python -m pytest tests/test_gpu_actor.py -q -x > gpurun_out/r2_pytest_actor.log 2>&1
tail -15 gpurun_out/r2_pytest_actor.log
for v in 0 1 2; do
  TACTOR_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.2 > gpurun_out/r2_bench_v$v.json 2> gpurun_out/r2_bench_v$v.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_v$v.json').readlines()[-1])
    print('variant $v', d['value'], d['stages'], d['roofline']['frac'], d['e2e']['ms_per_step'])
except Exception as e:
    print('variant $v failed', e); print(open('gpurun_out/r2_bench_v$v.err').read()[-2000:])
PY
done
for v in 0 1; do
  TACTOR_VARIANT=$v timeout 300 python bench.py --family large_bridge --batch 8192 --steps 20 --warmup 5 --cpu-seconds 0.2 > gpurun_out/r2_bench_large_v$v.json 2>/dev/null
  python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_large_v$v.json').readlines()[-1]);print('large variant $v', d['value'], d['stages'])"
done
