"""TEST INFRASTRUCTURE ONLY -- builds tests/cpu_emu/_build/libndp_emu.so: the product's kernel
sources compiled with g++ against the CUDA-on-CPU shim (cuda_emu.h).  Used by tests/test_emu_*.py
to check tiling / indexing / host orchestration against the oracle without a GPU.  The product
never loads this library."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "deformationpyramid_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libndp_emu.so")
SOURCES = ["ndp_warp_fwd.cu", "ndp_warp_bwd.cu", "ndp_warp_fwd_tc.cu", "ndp_warp_bwd_tc.cu", "ndp_warp_bwd_rc.cu", "ndp_chamfer.cu", "ndp_spatial.cu", "ndp_adam.cu", "ndp_cabi.cu"]


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "emu_hooks.cpp")] + \
           [os.path.join(ROOT, "include", "ndp_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-DNDP_EMU", "-ffp-contract=off", "-mfma", "-mf16c", "-pthread",
             "-Wno-unknown-pragmas", "-I", HERE, "-I", CSRC]

    def cc(src, lang_cuda):
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        cmd = ["g++"] + flags + (["-x", "c++"] if lang_cuda else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}{r.stderr}")
        return obj

    jobs = [(os.path.join(CSRC, s), True) for s in SOURCES] + \
           [(os.path.join(HERE, "cuda_emu.cpp"), False), (os.path.join(HERE, "emu_hooks.cpp"), False)]
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda j: cc(*j), jobs))
    r = subprocess.run(["g++", "-shared", "-pthread", "-o", LIB] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
