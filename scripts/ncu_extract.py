"""Summarise the ncu captures of a profile round (scripts/gpu_profile.sh) into profiles/:
  profiles/<tag>_full_<kernel>.raw.csv    ncu --page raw of the --set full capture
  profiles/<tag>_stalls_<kernel>.txt      hottest SASS lines / stall reasons (scripts/ncu_stalls.py)
  profiles/<tag>_launches_*.csv           the launch list (gpu__time_duration per launch) + per-kernel shares
  profiles/kernel_traffic.json            dram bytes / duration per launch per kernel (bench.py reads it)
Usage: python scripts/ncu_extract.py r01 [pairs_per_step] [stream_groups]"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 4          # stream groups of the solver (ndp_solver_cfg::streams)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PROF, exist_ok=True)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum"]
traffic = {"source": f"ncu --set full --clock-control none, bench.py --pairs {pairs} --iters 6 ({tag}): one launch per kernel "
                     f"= one stream group of {pairs // groups} pairs of 8192x8192 (cold-cache, serialised)",
           "pairs_per_launch": pairs // groups}
for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
    k = os.path.basename(rep)[5:-8]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, f"{tag}_full_{k}.raw.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, unit, val = rows[0], rows[1], rows[2]
    d = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            try:
                d[w] = float(val[i].replace(",", ""))
            except ValueError:
                d[w] = val[i]
            d[w + ".unit"] = unit[i]

    def to_bytes(name):
        v, u = d.get(name), d.get(name + ".unit", "byte")
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    d["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
    traffic[k] = d
    st = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls.py"), rep, "30"], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, f"{tag}_stalls_{k}.txt"), "w") as f:
        f.write(st)
with open(os.path.join(PROF, "kernel_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)

ll = os.path.join(OUT, "launches.csv")
if os.path.exists(ll):
    dst = os.path.join(PROF, f"{tag}_launches_bench_pairs{pairs}_iters6.csv")
    shutil.copy(ll, dst)
    rows = list(csv.reader(open(ll)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    kn, v, u = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) <= v:
            continue
        t = float(r[v].replace(",", ""))
        t = t / 1e3 if r[u] == "ns" else (t * 1e3 if r[u] == "ms" else t)
        agg[r[kn].split("(")[0]].append(t)
    tot = sum(sum(x) for x in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_shares_pairs{pairs}.txt"), "w") as f:
        f.write(f"# ncu launch list of `bench.py --steps 1 --warmup 1 --pairs {pairs} --iters 6` (per-launch times are cold-cache and serialised;\n"
                f"# each launch covers one stream group of {pairs // groups} pairs); shares of the summed kernel time\n")
        for n, x in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            line = f"{n[:60]:60s} n={len(x):5d} mean={sum(x) / len(x):9.2f} us  share={100 * sum(x) / tot:5.1f}%"
            print(line)
            f.write(line + "\n")
for name in ("bench_default.json", "bench_reference.json", "bench_pairs8.json", "bench_pairs16.json", "bench_fp32pipes_pairs16.json", "bench_2gpu.json",
             "smoke.log", "pytest_gpu.log", "sanitizer_memcheck.log", "sanitizer_racecheck.log", "sanitizer_synccheck.log", "host.txt", "nvidia-smi.txt"):
    src = os.path.join(OUT, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, f"{tag}_{name}" if not name.startswith(("host", "nvidia")) else name))
print(json.dumps({k: (v.get("dram_bytes_per_launch"), v.get("gpu__time_duration.sum")) for k, v in traffic.items() if isinstance(v, dict)}, indent=1))
