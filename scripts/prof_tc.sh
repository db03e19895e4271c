python bench.py --steps 3 --warmup 3 --cpu-seconds 0.1 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:actor_fused -s 4 -c 1 -o gpurun_out/prof_tc python bench.py --steps 3 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_tc.log 2>&1
echo rc=$?
