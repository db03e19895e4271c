# multi-GPU checks of round 2 (run under gpurun --gpus N): NCCL learner test (N >= 2), the bench at N ranks, the training leg
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_learner_nccl.py -x -q > gpurun_out/r2_pytest_nccl.log 2>&1; tail -3 gpurun_out/r2_pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --cpu-seconds 2 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_${N}gpu.json").readlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_gbs_per_rank"), d["e2e"].get("d2h_gbs_per_rank"))
for k,v in (d.get("extra") or {}).items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","envs_per_gpu","actor_ms","stages_ms","all_reduce","models_identical_across_ranks","updates_per_s")})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --train --steps 10 > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/r2_train_${N}gpu.err
tail -c 300 gpurun_out/r2_train_${N}gpu.err; cat gpurun_out/r2_train_${N}gpu.json
