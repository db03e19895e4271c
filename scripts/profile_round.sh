# Round profile capture (run under gpurun, 1 GPU).  Outputs go to gpurun_out/r2_*; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1 --no-extras"
$CMD > gpurun_out/r2_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -f -o gpurun_out/r2_prof_actor $CMD > gpurun_out/r2_ncu_actor.log 2>&1
echo actor rc=$?
ncu --set full --clock-control none --import-source on -k regex:tfem_step -s 12 -c 1 -f -o gpurun_out/r2_prof_fem $CMD > gpurun_out/r2_ncu_fem.log 2>&1
echo fem rc=$?
for k in actor fem; do
  ncu -i gpurun_out/r2_prof_$k.ncu-rep --page details > gpurun_out/r2_${k}_details.txt 2>&1
  ncu -i gpurun_out/r2_prof_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_raw.csv 2>&1
done
./mop_truss_marl_b200/lib/tensor_share > gpurun_out/r2_tensor_share.jsonl 2>&1
./mop_truss_marl_b200/lib/split_rate > gpurun_out/r2_split_rate.jsonl 2>&1
./mop_truss_marl_b200/lib/tfem_peaks > gpurun_out/r2_peaks.jsonl 2>&1
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 6 --warmup 1 --ref-step-seconds 1.0 > gpurun_out/r2_bench_reference.json 2>/dev/null
for cfg in "small_roof 16384" "large_bridge 8192" "large_roof 4096"; do set -- $cfg
  python bench.py --family $1 --batch $2 --steps 20 --warmup 5 --cpu-seconds 0.5 --no-extras > gpurun_out/r2_bench_actor_$1.json 2>/dev/null
done
python bench.py --train --steps 10 > gpurun_out/r2_train_1gpu.json 2>/dev/null
python scripts/dropin_episode.py > gpurun_out/r2_dropin_episode.json 2>&1
TACTOR_TRACE=gpurun_out/r2_trace.npy TFEM_LIB=mop_truss_marl_b200/lib/libtfem_prof.so python scripts/actor_prof.py small_bridge 4096 1 > gpurun_out/r2_actor_prof.jsonl 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
