OUT=gpurun_out; mkdir -p $OUT
for S in 4 6 8; do for P in 32 64; do
NDP_SOLVER_STREAMS=$S timeout 600 python bench.py --steps 1 --warmup 2 --pairs $P --no-cpu-baseline > $OUT/bench_s${S}_p$P.json 2> $OUT/bench_s${S}_p$P.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_s${S}_p$P.json")); print("streams=$S pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/bench_s${S}_p$P.err").read()[-800:])
PY
done; done
