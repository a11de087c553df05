#!/bin/bash
# ncu --set full captures (source-level stall samples) of the named kernels:  bash scripts/gpu_ncu.sh "k1 k2" [pairs]
OUT=gpurun_out; mkdir -p $OUT
P=${2:-8}
for K in $1; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 1 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
