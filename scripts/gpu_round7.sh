#!/bin/bash
# Status snapshot of the current tree: smoke, GPU parity, bench in both MLP modes, launch list, full
# ncu captures of the tensor-core MLP kernels and the pruned NN kernel.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -n "^E  \|Error\|^FAILED\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-250 | head -12
timeout 200 python scripts/phase_times.py 8 2>&1 | tail -2 | tee $OUT/phase_times.txt
summ() {
python - <<PY
import json
try:
    d=json.load(open("$1")); print("$1 value %.3f e2e %.3f"%(d["value"], d["e2e"]["value"]), {k: round(v,4) for k,v in d["kernel_ms_per_launch"].items()})
except Exception as e: print("$1 failed", e); print(open("$2").read()[-1500:])
PY
}
for P in 8 16; do
  timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_tc_p$P.json 2> $OUT/bench_tc_p$P.err
  summ $OUT/bench_tc_p$P.json $OUT/bench_tc_p$P.err
done
NDP_MLP_MODE=1 timeout 600 python bench.py --steps 1 --warmup 3 --pairs 16 --no-cpu-baseline > $OUT/bench_fp32_p16.json 2> $OUT/bench_fp32_p16.err
summ $OUT/bench_fp32_p16.json $OUT/bench_fp32_p16.err
echo "== default bench (with cpu baseline)"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
summ $OUT/bench_default.json $OUT/bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
cut -c1-400 $OUT/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches exit $?"
for K in ndp_warp_fwd_tc_kernel ndp_warp_bwd_tc_kernel ndp_nn_pruned_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
ls -la $OUT
