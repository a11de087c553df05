#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -n "^E  \|Error\|^FAILED\|passed\|failed" $OUT/pytest_gpu.log | cut -c1-250 | head -20
for K in ndp_warp_fwd_tc_kernel ndp_warp_bwd_tc_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs 8 --iters 6 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
