# A/B timing of TACTOR_FLAGS / TACTOR_VARIANT / library combinations: "flags:variant[:lib]" in $COMBOS
mkdir -p gpurun_out
for fv in ${COMBOS:-0:0 1:0}; do
  f=$(echo $fv | cut -d: -f1); v=$(echo $fv | cut -d: -f2); l=$(echo $fv | cut -d: -f3)
  lib=mop_truss_marl_b200/lib/libtfem${l:+_$l}.so
  TFEM_LIB=$lib TACTOR_FLAGS=$f TACTOR_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.05 > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_ab.json').readlines()[-1]);print('flags $f variant $v lib $l small', d['stages']['actor_ms'])
except Exception as e: print('flags $f variant $v failed', e, open('gpurun_out/r2_ab.err').read()[-800:])"
  if [ -n "$LARGE" ]; then
  TFEM_LIB=$lib TACTOR_FLAGS=$f TACTOR_VARIANT=$v timeout 300 python bench.py --family large_bridge --batch 8192 --steps 20 --warmup 5 --cpu-seconds 0.05 > gpurun_out/r2_ab.json 2>/dev/null
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_ab.json').readlines()[-1]);print('flags $f variant $v lib $l large', d['stages']['actor_ms'])
except Exception as e: print('large failed', e)"
  fi
done
