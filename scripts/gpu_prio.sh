#!/bin/bash
# per-launch priority experiment: NDP_DEBUG_PRIO="<tensor-core kernels>:<others>" (lower = served first)
OUT=gpurun_out; mkdir -p $OUT
python - <<PY
import torch
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")
PY
for V in "0:0" "-3:0" "0:-3" "-5:-1"; do
  NDP_DEBUG_PRIO=$V timeout 300 python bench.py --steps 2 --warmup 1 --pairs ${1:-32} --iters 60 --no-cpu-baseline --no-mode-b --no-config5 > $OUT/sw.json 2> $OUT/sw.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/sw.json")); print("prio $V: value*60/500 = %.2f"%(d["value"]*60/500.0), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/sw.err").read()[-600:])
PY
done
