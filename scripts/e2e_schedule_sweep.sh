# end-to-end step against the piece schedule (run under gpurun, 1 GPU): one bench line per schedule -> gpurun_out/r2_e2e_schedules.jsonl
mkdir -p gpurun_out; : > gpurun_out/r2_e2e_schedules.jsonl
for s in ${SCHEDULES:-2 3 512,3584 1024,3072 512,1536,2048 256,1280,2560 512,1024,2560 256,768,3072 256,768,1024,2048 1024,1024,2048 128,896,3072}; do
  python bench.py --steps 30 --warmup 5 --cpu-seconds 0.1 --no-extras --e2e-pieces $s 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); e = d['e2e']
print(json.dumps({'pieces': '$s', 'e2e': e['value'], 'ms_per_4096': 4096 / e['value'] * 1e3, 'compact': d.get('e2e_compact_state', {}).get('value')}))" | tee -a gpurun_out/r2_e2e_schedules.jsonl
done
