"""ORACLE - TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Builds oracle/_build/libndp_oracle.so from oracle/knn_oracle.c with gcc.

The reference (/root/reference) is pure Python on this path, so there is nothing to compile
into oracle/_ref/: the Python reference is imported in the build container by
oracle/gen_golden.py to pin the restatement, and its outputs travel as tests/golden/*.npz.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libndp_oracle.so")


def _cpu_has(flag: str) -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return flag in line.split()
    except OSError:
        pass
    return False


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "knn_oracle.c")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
           "-o", LIB, src, "-lm"]
    # A hardware FMA makes fmaf() one instruction; without it libm's (correctly rounded,
    # slow) software fmaf is used -- the results are identical either way.
    # NB: the .so built in the build container travels to the GPU box, so only enable it when
    # explicitly asked for (the box CPU is not known here).
    if os.environ.get("NDP_ORACLE_MFMA", "1") == "1" and _cpu_has("fma"):
        cmd.insert(1, "-mfma")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
