OUT=gpurun_out; mkdir -p $OUT
for T in 4 8 16; do for P in 32 8; do
NDP_BWD_TPC=$T timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_tpc$T.json 2> $OUT/bench_tpc$T.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_tpc$T.json")); print("tpc=$T pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/bench_tpc$T.err").read()[-800:])
PY
done; done
