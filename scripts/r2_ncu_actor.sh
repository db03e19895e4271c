# ncu --set full capture of the actor kernel (after the plain run exited 0); TAG names the outputs, TFEM_LIB picks the build
mkdir -p gpurun_out
TAG=${TAG:-actor}
CMD="python bench.py --steps 4 --warmup 3 --cpu-seconds 0.1"
$CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:actor_pipe -s 4 -c 1 -f -o gpurun_out/r2_prof_$TAG $CMD > gpurun_out/r2_ncu_$TAG.log 2>&1
echo ncu rc=$?
ncu -i gpurun_out/r2_prof_$TAG.ncu-rep --page details > gpurun_out/r2_${TAG}_details.txt 2>&1
ncu -i gpurun_out/r2_prof_$TAG.ncu-rep --page raw --csv > gpurun_out/r2_${TAG}_raw.csv 2>&1
ncu -i gpurun_out/r2_prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_${TAG}_source.csv 2>&1
