"""Aggregate the per-instruction stall samples of an ncu report (source page) and list the hottest
SASS lines.  Usage: python scripts/ncu_stalls.py gpurun_out/prof_X.ncu-rep [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
# possibly several kernels: take the first
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and i > start]
body = lines[start:(end[0] if end else len(lines))]
rows = list(csv.DictReader(io.StringIO("\n".join(body))))
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter()
for r in rows:
    for k in stalls:
        tot[k] += int(r[k] or 0)
allsamp = sum(tot.values())
print("total samples", allsamp)
for k, v in tot.most_common(10):
    print(f"  {k:28s} {v:8d} {100.0*v/allsamp:5.1f}%")
print("instructions executed:", sum(int(r["Instructions Executed"] or 0) for r in rows))
ops = collections.Counter()
for r in rows:
    op = r["Source"].split()[0] if r["Source"].split() else "?"
    if op.startswith("@"): op = r["Source"].split()[1]
    ops[op.split(".")[0]] += int(r["Instructions Executed"] or 0)
print("op mix:", [(k, v) for k, v in ops.most_common(14)])
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:top]:
    s = {k: int(r[k] or 0) for k in stalls if int(r[k] or 0)}
    main = sorted(s.items(), key=lambda kv: -kv[1])[:3]
    print(f"{int(r['# Samples']):7d}  {r['Source'][:70]:70s} {main}")
