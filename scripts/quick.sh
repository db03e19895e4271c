python scripts/dbg_actor_err.py 2>&1 | grep -v Warn | grep -v "return 1.0"
python bench.py --steps 20 --warmup 3 --cpu-seconds 0.1 > gpurun_out/plain3.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/plain3.log').readlines()[-1]);print(d['value'], d['stages'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
