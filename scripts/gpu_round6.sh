#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -n "^E  \|Error\|^FAILED\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-250 | head -12
timeout 200 python scripts/phase_times.py 8 2>&1 | tail -2
for P in 8 16; do
  timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_tc_p$P.json 2> $OUT/bench_tc_p$P.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_tc_p$P.json")); print("pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_iteration"].items()})
except Exception as e: print("pairs=$P failed", e); print(open("$OUT/bench_tc_p$P.err").read()[-1500:])
PY
done
