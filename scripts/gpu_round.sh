#!/bin/bash
# One gpurun call: parity tests, smoke, short bench, sanitizer, ncu launch list + full captures.
# Everything lands in gpurun_out/.  Usage (from the repo root, on the GPU box):
#   bash scripts/gpu_round.sh [quick|full]
MODE=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" > $OUT/host.txt 2>&1
lscpu | head -20 >> $OUT/host.txt 2>&1

echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -40 $OUT/pytest_gpu.log

echo "== smoke"
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -5 $OUT/smoke.log

echo "== bench (short)"
timeout 900 python bench.py --steps 2 --warmup 3 --pairs 8 > $OUT/bench_short.json 2> $OUT/bench_short.err
echo "bench exit $?"; tail -c 3000 $OUT/bench_short.json; tail -5 $OUT/bench_short.err

if [ "$MODE" = "full" ]; then
  echo "== bench pairs sweep"
  for P in 1 2 4 16; do
    timeout 600 python bench.py --steps 1 --warmup 3 --pairs $P --no-cpu-baseline > $OUT/bench_p$P.json 2> $OUT/bench_p$P.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_p$P.json")); print("pairs=$P value", d["value"], "e2e", d["e2e"]["value"], d["kernel_ms_per_iteration"])
except Exception as e: print("pairs=$P failed", e)
PY
  done
  echo "== compute-sanitizer (memcheck) on the smoke run"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > $OUT/sanitizer_memcheck.log 2>&1
  echo "memcheck exit $?"; tail -4 $OUT/sanitizer_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py --smoke > $OUT/sanitizer_racecheck.log 2>&1
  echo "racecheck exit $?"; tail -4 $OUT/sanitizer_racecheck.log

  echo "== ncu launch list (the bench command, reduced iterations)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  echo "== ncu full captures"
  for K in ndp_nn_kernel ndp_warp_fwd_kernel ndp_warp_bwd_kernel ndp_chamfer_reduce_kernel ndp_reduce_adam_kernel; do
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 2 -f -o $OUT/prof_$K \
        python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1
    echo "ncu $K exit $?"
  done
fi
ls -la $OUT
