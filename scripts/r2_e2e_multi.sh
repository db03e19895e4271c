# end-to-end step on N GPUs: transfer kernels against copy engines (run under gpurun --gpus N)
N=${N:-2}
mkdir -p gpurun_out; : > gpurun_out/r2_e2e_multi_${N}gpu.jsonl
run() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-extras --cpu-seconds 0.2 "$@" > gpurun_out/_e2e_multi.out 2>/dev/null
  grep '^{' gpurun_out/_e2e_multi.out | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(json.dumps({'n_gpus': $N, 'mode': '$tag', 'e2e': d['e2e']['value'], 'ms': d['e2e']['ms_per_step'], 'compact': d.get('e2e_compact_state', {}).get('value'), 'value': d['value']}))" | tee -a gpurun_out/r2_e2e_multi_${N}gpu.jsonl
}
run kernels-auto
run kernels-2 --e2e-pieces 2
run kernels-4 --e2e-pieces 4
export TROLLOUT_ZEROCOPY=0
run engines-2 --e2e-pieces 2
run engines-4 --e2e-pieces 4
