#!/bin/bash
# ncu --set full capture of the named kernels (+ launch list):  bash scripts/gpu_ncu2.sh "k1 k2" [pairs]
OUT=gpurun_out; mkdir -p $OUT
P=${2:-32}
rm -f $OUT/prof_*.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline --no-mode-b > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
for K in $1; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline --no-mode-b > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
