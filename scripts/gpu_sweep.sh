OUT=gpurun_out; mkdir -p $OUT
NDP_FWD_ROUNDS2=2 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
for R in 1 2 4; do for P in 8 32; do
NDP_FWD_ROUNDS2=$R timeout 600 python bench.py --steps 1 --warmup 2 --pairs $P --no-cpu-baseline > $OUT/bench_sw.json 2> $OUT/bench_sw.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_sw.json")); print("rounds2=$R pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), round(d["kernel_ms_per_launch"]["warp_fwd"],4))
except Exception as e: print("failed", e); print(open("$OUT/bench_sw.err").read()[-500:])
PY
done; done
