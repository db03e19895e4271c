# GPU call 2 of round 2: pipe peaks, host-pipeline + actor tests, ncu of the actor kernel, bench
mkdir -p gpurun_out
./mop_truss_marl_b200/lib/tfem_peaks > gpurun_out/r2_peaks.jsonl 2>&1; tail -13 gpurun_out/r2_peaks.jsonl
timeout 1500 python -m pytest tests/test_gpu_actor.py tests/test_gpu_host_pipeline.py -q -x > gpurun_out/r2_pytest2.log 2>&1
tail -8 gpurun_out/r2_pytest2.log
CMD="python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1"
timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.2 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench2.json').readlines()[-1]);print(d['value'], d['stages'], d['roofline']['frac'], d['e2e'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -f -o gpurun_out/r2_prof_actor $CMD > gpurun_out/r2_ncu_actor.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_prof_actor.ncu-rep --page details > gpurun_out/r2_actor_details.txt 2>&1
ncu -i gpurun_out/r2_prof_actor.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_actor_source.csv 2>&1
ls -la gpurun_out/r2_prof_actor.ncu-rep
