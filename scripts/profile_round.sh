# Round profile capture (run under gpurun, 1 GPU).  Outputs go to gpurun_out/.
set -x
CMD="python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1"
$CMD > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -o gpurun_out/prof_actor $CMD > gpurun_out/ncu_actor.log 2>&1
echo actor rc=$?
ncu --set full --clock-control none --import-source on -k regex:tfem_step -s 12 -c 1 -o gpurun_out/prof_fem $CMD > gpurun_out/ncu_fem.log 2>&1
echo fem rc=$?
