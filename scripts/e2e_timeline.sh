#!/bin/bash
# per-piece event times of the end-to-end rollout step (TROLLOUT_TIMELINE=1: direct enqueue with timing events)
for p in 1 2 4 8; do
  echo "== pieces $p"
  TROLLOUT_TIMELINE=1 python bench.py --steps 5 --warmup 3 --cpu-seconds 0.05 --e2e-pieces $p 2>&1 >/dev/null | grep "trollout timeline" | tail -2
done 2>&1 | tee gpurun_out/e2e_timeline.txt
