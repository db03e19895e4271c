#!/bin/bash
# the bench.py lines kept under profiles/: default workload, FEM-only and actor + FEM for the other families
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err
for cfg in "small_roof 16384" "large_bridge 8192" "large_roof 4096"; do set -- $cfg
  python bench.py --family $1 --batch $2 --no-actor --steps 50 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_$1.json 2>/dev/null
  python bench.py --family $1 --batch $2 --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_actor_$1.json 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pareto|actor_pipe|tfem_step" -s 30 -c 9 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -f -o gpurun_out/prof_actor python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_actor.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
