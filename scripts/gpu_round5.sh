#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
for P in 8 16; do
  timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_tc_p$P.json 2> $OUT/bench_tc_p$P.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_tc_p$P.json")); print("pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_iteration"].items()})
except Exception as e: print("pairs=$P failed", e); print(open("$OUT/bench_tc_p$P.err").read()[-1500:])
PY
done
for K in ndp_warp_fwd_tc_kernel ndp_warp_bwd_tc_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
