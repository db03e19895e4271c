# A/B of the actor kernel variants (run under gpurun): tests, then bench with the CTA-pair and single-CTA builds
timeout 300 python -m pytest tests/test_gpu_actor.py -x -q 2>&1 | tail -5 | cut -c1-300
for n in 2 1; do
  TACTOR_NCTA=$n timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.1 2>&1 | python -c "
import sys,json
l=sys.stdin.readlines()
try:
    d=json.loads(l[-1]); print('ncta=$n value %.4e ms %.4f actor_ms %.4f fem_ms %.4f'%(d['value'],d['ms_per_step'],d['stages']['actor_ms'],d['stages']['fem_ms']))
except Exception as e:
    print('ncta=$n FAILED', ''.join(l[-8:])[:1500])
"
done
