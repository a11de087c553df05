"""Per-kernel SASS instruction summary of the built library (profiles/<tag>_sass_summary.txt): the mnemonics that prove
the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk = TMA), SYNCS (mbarrier), plus the FP32-pipe and shared-memory instruction counts.
Usage: python scripts/sass_summary.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "deformationpyramid_b200", "lib", "libndp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "SYNCS", "FFMA", "FADD", "F2FP", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "ATOMG", "RED", "SHFL"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
lines = [f"# cuobjdump -sass of deformationpyramid_b200/lib/libndp_b200.so (sm_100a), instruction counts per kernel ({tag})",
         "# kernel".ljust(48) + "total " + " ".join(w.rjust(7) for w in WANT)]
for k, c in counts.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    lines.append(name[:46].ljust(48) + str(total[k]).rjust(5) + " " + " ".join(str(c.get(w, 0)).rjust(7) for w in WANT))
txt = "\n".join(lines) + "\n"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w") as f:
    f.write(txt)
print(txt)
