python scripts/dbg_actor_err.py 2>&1 | grep -v Warn | grep -v "return 1.0"
python bench.py --steps 20 --warmup 3 --cpu-seconds 0.1 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -o gpurun_out/prof_pipe -f python bench.py --steps 3 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_pipe.log 2>&1
echo rc=$?
tail -1 gpurun_out/plain3.log | cut -c1-200
