#!/bin/bash
# Round profile: parity tests, smoke, bench (default + reference arm + pair sweep), sanitizer, ncu launch
# list and full captures of every kernel of the iteration.  Everything lands in gpurun_out/.
OUT=gpurun_out; mkdir -p $OUT
P=${1:-32}
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" > $OUT/host.txt 2>&1; lscpu | head -20 >> $OUT/host.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -1 $OUT/pytest_gpu.log
timeout 1200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"; cut -c1-300 $OUT/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cut -c1-200 $OUT/bench_reference.json
for Q in 8 16; do
  timeout 600 python bench.py --steps 2 --warmup 3 --pairs $Q --no-cpu-baseline --no-mode-b > $OUT/bench_pairs$Q.json 2> $OUT/bench_pairs$Q.err; echo "bench pairs $Q exit $?"
done
timeout 900 python bench.py --mlp fp32 --steps 1 --warmup 3 --pairs 16 --no-cpu-baseline --no-mode-b > $OUT/bench_fp32pipes_pairs16.json 2> $OUT/bench_fp32pipes.err; echo "fp32-pipe bench exit $?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -2 $OUT/sanitizer_memcheck.log
for T in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $T --error-exitcode 7 python __graft_entry__.py --smoke > $OUT/sanitizer_$T.log 2>&1; echo "$T exit $?"; tail -2 $OUT/sanitizer_$T.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline --no-mode-b > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
for K in ndp_warp_bwd_rc_kernel ndp_warp_fwd_tc2_kernel ndp_nn_pruned_kernel ndp_reduce_adam_kernel ndp_head_grad_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 1 -f -o $OUT/prof_$K \
      python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline --no-mode-b > $OUT/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
ls $OUT | wc -l
