"""TEST INFRASTRUCTURE: runs one of the reference's CLI scripts UNMODIFIED (runpy, __main__) against the mirror
modules (deformationpyramid_b200.install_as_model) with the headless open3d / easydict stand-ins, in a process of
its own.  On a GPU-less machine the CPU-emulated build of the kernel sources (tests/cpu_emu) stands in for the
device (the product never loads it) and torch.cuda.current_device is pointed at the CPU, because
shape_transfer.py:28,53 hard-codes gpu_mode.

    python tests/ref_driver.py <reference root> <script> <workdir> <out.npz> '<json options>'
options: argv (list), emu (bool), overrides (dict applied to every EasyDict the script builds: the harness's way of
shrinking shape_transfer.py's hard-coded workload), seed (surface-sampling seed of the open3d stand-in)."""
import glob
import json
import os
import runpy
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ref, script, workdir, out, opts = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], json.loads(sys.argv[5])
    import deformationpyramid_b200 as ndp
    from deformationpyramid_b200 import _lib
    if opts.get("emu"):
        from emu_util import emu_lib
        _lib._LIB = emu_lib()
        torch.cuda.current_device = lambda: torch.device("cpu")
    ndp.headless.install(force=True, seed=int(opts.get("seed", 0)))
    overrides = opts.get("overrides") or {}
    if overrides:
        import easydict
        base = easydict.EasyDict

        class Shrunk(base):
            def __init__(self, d=None, **kw):
                super().__init__(d, **kw)
                for k, v in overrides.items():
                    if k in self:
                        self[k] = v
        easydict.EasyDict = Shrunk
    sys.path.insert(0, ref)
    os.chdir(workdir)
    ndp.install_as_model()
    sys.argv = [script] + list(opts.get("argv", []))
    g = runpy.run_path(os.path.join(ref, script), run_name="__main__")
    res = {"modules": np.array([sys.modules["model.registration"].__file__, sys.modules["model.nets"].__file__,
                                sys.modules["model.loss"].__file__, getattr(sys.modules.get("model.geometry"), "__file__", "")])}
    if script == "shape_transfer.py":
        res.update(warped_vert=np.asarray(g["warped_vert"], np.float32),
                   src_mean=g["src_mean"].cpu().numpy(), tgt_mean=g["tgt_mean"].cpu().numpy(),
                   src_pcd=g["src_pcd"].cpu().numpy(), tgt_pcd=g["tgt_pcd"].cpu().numpy(),
                   mesh_vert=g["mesh_vert"].cpu().numpy(), drawn=np.array([ndp.headless._state["drawn"]]))
    else:
        cfg = g["config"]
        for split in g["splits"]:
            res["entries_" + split] = np.array(glob.glob(os.path.join(cfg.data_root, split, "*/*.npz")))
            with open(os.path.join(cfg.snapshot_dir, split + ".log")) as f:
                res["log_" + split] = np.array([f.read()])
        res["timer_keys"] = np.array(sorted(g["timer"].timers.keys()))
    np.savez(out, **res)


if __name__ == "__main__":
    main()
