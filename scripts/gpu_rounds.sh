OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -1 $OUT/pytest_gpu.log
for T in 1 2 4; do for P in 32 8; do
NDP_FWD_ROUNDS=$T timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_r$T.json 2> $OUT/bench_r$T.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_r$T.json")); print("rounds=$T pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/bench_r$T.err").read()[-800:])
PY
done; done
