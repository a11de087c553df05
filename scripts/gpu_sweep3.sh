#!/bin/bash
# pairs-per-step x stream-groups sweep at the headline shape with reduced iterations: bash scripts/gpu_sweep3.sh "<pairs:streams ...>" [iters] [tpc] [rounds]
OUT=gpurun_out; mkdir -p $OUT
IT=${2:-60}; T=${3:-8}; R=${4:-4}
for PS in $1; do P=${PS%%:*}; S=${PS##*:}
  timeout 300 python bench.py --steps 2 --warmup 1 --pairs $P --iters $IT --no-cpu-baseline --no-mode-b --no-config5 --streams $S --tpc $T --fwd-rounds $R > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("pairs=$P streams=$S tpc=$T rounds=$R: value*iters/500 = %.2f pairs/s-equivalent"%(d["value"]*$IT/500.0), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("pairs=$P streams=$S failed", e); print(open("$OUT/sw.err").read()[-800:])
PY
done
