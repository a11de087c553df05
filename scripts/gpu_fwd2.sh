OUT=gpurun_out; mkdir -p $OUT
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -n "^E  \|Error\|^FAILED\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-250 | head -8
timeout 200 python scripts/phase_times.py 8 2>&1 | tail -2
for V in 2 1; do for P in 32 8; do
NDP_FWD_TC_VERSION=$V timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_v$V.json 2> $OUT/bench_v$V.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$V.json")); print("fwd_version=$V pairs=$P value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("failed", e); print(open("$OUT/bench_v$V.err").read()[-800:])
PY
done; done
