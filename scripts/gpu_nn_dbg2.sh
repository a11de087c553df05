#!/bin/bash
# in-step cost of the NN kernel's parts: headline shape, 32 pairs / 4 groups, with NDP_DEBUG_NN / NDP_DEBUG_SKIP variants
OUT=gpurun_out; mkdir -p $OUT
run() { env $1 timeout 300 python bench.py --steps 2 --warmup 1 --pairs 32 --iters 60 --no-cpu-baseline --no-mode-b --no-config5 > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("$1: ms/step*500/60 = %.1f"%(d["ms_per_step"]*500.0/60), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/sw.err").read()[-600:])
PY
}
for V in "NDP_DEBUG_NN=0" "NDP_DEBUG_NN=1" "NDP_DEBUG_SKIP=2"; do run $V; done
