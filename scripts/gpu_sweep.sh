OUT=gpurun_out; mkdir -p $OUT
run() {
NDP_SOLVER_STREAMS=$1 timeout 600 python bench.py --steps 1 --warmup 2 --pairs $2 --no-cpu-baseline > $OUT/bench_sw.json 2> $OUT/bench_sw.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_sw.json")); print("streams=$1 pairs=$2 value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]))
except Exception as e: print("failed", e); print(open("$OUT/bench_sw.err").read()[-500:])
PY
}
run 4 32; run 4 40; run 4 48; run 4 64; run 3 32; run 5 32; run 5 40; run 6 48; run 8 64
