#!/bin/bash
# Execution-profile sweep at the headline shape with reduced iterations:  bash scripts/gpu_sweep2.sh "<streams list>" "<tpc list>" "<rounds list>" [pairs] [iters]
OUT=gpurun_out; mkdir -p $OUT
P=${4:-32}; IT=${5:-60}
for S in $1; do for T in $2; do for R in $3; do
  timeout 300 python bench.py --steps 2 --warmup 1 --pairs $P --iters $IT --no-cpu-baseline --no-mode-b --no-config5 --streams $S --tpc $T --fwd-rounds $R > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("streams=$S tpc=$T rounds=$R pairs=$P: value*iters/500 = %.2f pairs/s-equivalent"%(d["value"]*$IT/500.0), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("streams=$S tpc=$T rounds=$R failed", e); print(open("$OUT/sw.err").read()[-800:])
PY
done; done; done
