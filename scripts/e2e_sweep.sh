#!/bin/bash
# end-to-end (host buffers) step time against the number of pieces the batch is cut into
mkdir -p gpurun_out
for p in 1 2 3 4 6 8 16; do
  python bench.py --steps 40 --warmup 5 --cpu-seconds 0.1 --e2e-pieces $p 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1]); e = d['e2e']
print(json.dumps({'pieces': $p, 'e2e_ms': e['ms_per_step'], 'e2e_value': e['value'], 'value': d['value']}))"
done | tee gpurun_out/e2e_sweep.jsonl
