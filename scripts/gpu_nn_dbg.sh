#!/bin/bash
# isolated NN launch time with parts of the kernel switched off (NDP_DEBUG_NN), one stream group of 8 pairs
OUT=gpurun_out; mkdir -p $OUT
for D in ${1:-0 1 2 4 6 7}; do
  NDP_DEBUG_NN=$D timeout 300 python bench.py --steps 2 --warmup 1 --pairs 8 --iters 60 --streams 1 --no-cpu-baseline --no-mode-b --no-config5 > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("NDP_DEBUG_NN=$D:", {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/sw.err").read()[-600:])
PY
done
