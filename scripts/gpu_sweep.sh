OUT=gpurun_out; mkdir -p $OUT
for P in 24 32 48 64; do
timeout 600 python bench.py --steps 1 --warmup 2 --pairs $P --no-cpu-baseline > $OUT/bench_sw.json 2> $OUT/bench_sw.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_sw.json")); print("pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]))
except Exception as e: print("failed", e); print(open("$OUT/bench_sw.err").read()[-500:])
PY
done
