#!/bin/bash
# Marginal in-step cost of each kernel: the headline shape with one kernel class left out at a time (NDP_DEBUG_SKIP; results are garbage,
# only the step time is read).  bash scripts/gpu_marginal.sh [pairs] [iters]
OUT=gpurun_out; mkdir -p $OUT
P=${1:-32}; IT=${2:-60}
for SK in 0 1 2 4 8 6 14 7 ; do
  NDP_DEBUG_SKIP=$SK timeout 300 python bench.py --steps 2 --warmup 1 --pairs $P --iters $IT --no-cpu-baseline --no-mode-b --no-config5 > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("skip=$SK pairs=$P: ms/step*500/iters = %.1f"%(d["ms_per_step"]*500.0/$IT), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("skip=$SK failed", e); print(open("$OUT/sw.err").read()[-600:])
PY
done
