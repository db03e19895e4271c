# A/B of the actor kernel variants (run under gpurun): tests, then bench per variant
timeout 300 python -m pytest tests/test_gpu_actor.py -x -q 2>&1 | tail -3 | cut -c1-300
for cfg in "pipe 1" "pipe 2" "fused 1"; do set -- $cfg
  TACTOR_KERNEL=$1 TACTOR_NCTA=$2 timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 0.1 2>&1 | python -c "
import sys,json
l=sys.stdin.readlines()
try:
    d=json.loads(l[-1]); print('$1 ncta=$2 value %.4e ms %.4f actor_ms %.4f fem_ms %.4f'%(d['value'],d['ms_per_step'],d['stages']['actor_ms'],d['stages']['fem_ms']))
except Exception as e:
    print('$1 ncta=$2 FAILED', ''.join(l[-8:])[:1500])
"
done
