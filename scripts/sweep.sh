for v in v218 v328; do
  export TFEM_LIB=$PWD/mop_truss_marl_b200/lib/libtfem_$v.so
  if [ "$v" = "v218" ]; then timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | cut -c1-200; fi
  for cfg in "small_bridge 4096" "small_roof 16384" "large_bridge 8192"; do set -- $cfg
    python bench.py --family $1 --batch $2 --steps 50 --warmup 5 --cpu-seconds 0.2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v', d['config']['family'], d['config']['envs_per_gpu'], 'value %.3e'%d['value'], 'ms %.4f'%d['ms_per_step'])"
  done
done
