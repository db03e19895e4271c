# round-2 iteration check (1 GPU): actor tests, per-warp cycle counters + time line, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_actor.py -q -x > gpurun_out/r2_pytest_actor.log 2>&1; tail -5 gpurun_out/r2_pytest_actor.log
TACTOR_TRACE=gpurun_out/r2_trace.npy TFEM_LIB=mop_truss_marl_b200/lib/libtfem_prof.so python scripts/actor_prof.py small_bridge 4096 1 > gpurun_out/r2_actor_prof.jsonl 2>&1; grep -v "\"warp\": \(1\|2\|3\|5\|6\|7\|9\|10\|11\|13\|14\|15\)," gpurun_out/r2_actor_prof.jsonl | cut -c1-250
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.2 > gpurun_out/r2_bench_iter.json 2> gpurun_out/r2_bench_iter.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_iter.json').readlines()[-1]);print(d['value'], d['stages'], d['roofline']['frac'], d['e2e']['value'], d['status_nonzero_envs'])"
