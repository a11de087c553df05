#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for MODE in 0 1; do
  for P in 1 8 16; do
    timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --nn-mode $MODE --no-cpu-baseline > $OUT/bench_m${MODE}_p$P.json 2> $OUT/bench_m${MODE}_p$P.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_m${MODE}_p$P.json")); print("mode=$MODE pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_iteration"].items()})
except Exception as e: print("mode=$MODE pairs=$P failed", e); print(open("$OUT/bench_m${MODE}_p$P.err").read()[-2000:])
PY
  done
done
echo "== ncu pruned kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ndp_nn_pruned_kernel -s 40 -c 2 -f -o $OUT/prof_ndp_nn_pruned_kernel \
    python bench.py --steps 1 --warmup 1 --pairs 8 --iters 30 --no-cpu-baseline > $OUT/ncu_pruned.log 2>&1
echo "ncu exit $?"
ls -la $OUT | tail -12
