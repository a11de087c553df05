#!/bin/bash
# Quick GPU check: smoke, GPU parity tests, phase stamps, bench at the given pair counts.
#   bash scripts/gpu_quick.sh "8 16" [tag]
OUT=gpurun_out; mkdir -p $OUT
PAIRS=${1:-"8 16"}; TAG=${2:-q}
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -n "^E  \|Error\|^FAILED\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-250 | head -12
timeout 200 python scripts/phase_times.py 8 2>&1 | tail -2 | tee $OUT/phase_times_$TAG.txt
for P in $PAIRS; do
  timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_${TAG}_p$P.json 2> $OUT/bench_${TAG}_p$P.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${TAG}_p$P.json")); print("pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("pairs=$P failed", e); print(open("$OUT/bench_${TAG}_p$P.err").read()[-1500:])
PY
done
