# usage: bash scripts/prof_variant.sh <lib variant> <family> <batch> <outname>
export TFEM_LIB=$PWD/mop_truss_marl_b200/lib/libtfem_$1.so
python bench.py --family $2 --batch $3 --steps 10 --warmup 3 --cpu-seconds 0.1 > gpurun_out/plain_$4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tfem_step -s 15 -c 1 -o gpurun_out/$4 python bench.py --family $2 --batch $3 --steps 10 --warmup 3 --cpu-seconds 0.1 > gpurun_out/ncu_$4.log 2>&1
echo rc=$?
