# Round profile capture (run under gpurun, 1 GPU).  Outputs go to gpurun_out/; summaries are copied to profiles/ by hand.
set -x
CMD="python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -f -o gpurun_out/prof_actor $CMD > gpurun_out/ncu_actor.log 2>&1
echo actor rc=$?
ncu --set full --clock-control none --import-source on -k regex:tfem_step -s 12 -c 1 -f -o gpurun_out/prof_fem $CMD > gpurun_out/ncu_fem.log 2>&1
echo fem rc=$?
python scripts/solver_compare.py large_bridge 8192 > gpurun_out/solver_compare_large.json 2>&1 && \
ncu --set full --clock-control none -k regex:dense_dmma -s 3 -c 1 -f -o gpurun_out/prof_dense python scripts/solver_compare.py large_bridge 8192 > gpurun_out/ncu_dense.log 2>&1
echo dense rc=$?
ncu --set full --clock-control none -k regex:tfem_step -s 3 -c 1 -f -o gpurun_out/prof_banded python scripts/solver_compare.py large_bridge 8192 > gpurun_out/ncu_banded.log 2>&1
echo banded rc=$?
python scripts/solver_compare.py small_bridge 8192 > gpurun_out/solver_compare_small.json 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err
for cfg in "small_roof 16384" "large_bridge 8192" "large_roof 4096"; do set -- $cfg
  python bench.py --family $1 --batch $2 --no-actor --steps 50 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_$1.json 2>/dev/null
  python bench.py --family $1 --batch $2 --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_actor_$1.json 2>/dev/null
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
