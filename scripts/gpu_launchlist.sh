#!/bin/bash
# ncu launch list (per-launch durations) of a reduced-iteration bench run -> gpurun_out/launches.csv
OUT=gpurun_out; mkdir -p $OUT
P=${1:-8}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --pairs $P --iters 6 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches exit $?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; k = H.index("Kernel Name"); v = H.index("Metric Value"); u = H.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) <= v: continue
    t = float(r[v].replace(",", "")); 
    if r[u] == "ns": t /= 1e3
    elif r[u] == "ms": t *= 1e3
    agg[r[k].split("(")[0]].append(t)
tot = sum(sum(x) for x in agg.values())
for n, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{n:36s} n={len(x):4d} mean={sum(x)/len(x):8.2f} us  share={100*sum(x)/tot:5.1f}%")
PY
