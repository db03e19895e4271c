for b in 1184 2368 3552 4096 4144 4736; do
  python bench.py --batch $b --steps 40 --warmup 5 --cpu-seconds 0.05 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('B', $b, 'actor_ms', round(d['stages']['actor_ms'],4), 'fem', round(d['stages'].get('fem_ms', d['stages'].get('env_ms',0)),4))"
done
TACTOR_NO_SPLIT=1 python bench.py --batch 4096 --steps 40 --warmup 5 --cpu-seconds 0.05 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('nosplit 4096 actor_ms', round(d['stages']['actor_ms'],4))"
