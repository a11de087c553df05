"""Row (b) of SURVEY.md section 8: the reference's CLI scripts drop in UNCHANGED.

Runs /root/reference/eval_nolearned.py and /root/reference/shape_transfer.py byte for byte as they are (runpy,
__main__, in a process of their own -- tests/ref_driver.py) against the mirror modules installed by
deformationpyramid_b200.install_as_model(), with the package's headless open3d / easydict stand-ins, on a synthetic
two-split 4DMatch directory (.npz wire format of correspondence/datasets/_4dmatch.py:60-73) and two small PLY
meshes.  There is no GPU in the build container, so the CPU-emulated build of the SAME kernel sources stands in for
the device; what the scripts print / return is compared with the oracle replaying the same RNG stream.
The reference checkout does not travel to the GPU box: there these tests skip, and tests/test_gpu_api.py runs the
package's own equivalents (shard.evaluate, shape_transfer.shape_transfer) on the hardware.
"""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "eval_nolearned.py")),
                                reason="the reference checkout is only present in the build container")

from oracle import ndp_oracle as O  # noqa: E402
from deformationpyramid_b200.synthetic import make_pair  # noqa: E402


def _drive(script, workdir, out, **opts):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_driver.py"), REF, script, str(workdir), str(out), json.dumps(opts)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return dict(np.load(out, allow_pickle=False)), r.stdout


def _write_4dmatch(root, split, seqs, n, seed0):
    """data_root/split/<seq>/<pair>.npz with the keys _4dmatch.py:60-73 reads."""
    k = seed0
    for s in range(seqs):
        d = os.path.join(root, split, f"seq{s:03d}")
        os.makedirs(d)
        for p in range(2):
            src, tgt = make_pair(k, n + 7 * p, n - 5 * p)
            g = np.random.default_rng(k)
            ang = 0.1
            rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
            trans = g.normal(0, 0.02, (3, 1)).astype(np.float32)
            flow = (0.02 * np.sin(3.0 * src.numpy()[:, [1, 2, 0]])).astype(np.float32)
            corr = np.stack([np.arange(0, src.shape[0], 2), np.arange(0, src.shape[0], 2) % tgt.shape[0]], 1)
            np.savez(os.path.join(d, f"cam1_{p:04d}_cam2_{p + 1:04d}.npz"), rot=rot, trans=trans, s2t_flow=flow,
                     s_pc=src.numpy(), t_pc=tgt.numpy(), correspondences=corr)
            k += 1


CFG = dict(iters=2, lr=0.01, max_break_count=15, break_threshold_ratio=0.001, w_reg=0.0, samples=96, m=2, k0=-8,
           depth=3, width=128, motion_type="SE3", rotation_format="axis_angle")


def test_eval_nolearned_runs_unmodified_against_the_mirror(tmp_path):
    data = tmp_path / "data"
    _write_4dmatch(str(data), "4DMatch-F", 1, 140, 300)
    _write_4dmatch(str(data), "4DLoMatch-F", 1, 130, 400)
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text("gpu_mode: False\ndeformation_model: NDP\nuse_ldmk: False\nuse_depth: False\n"
                   + "".join(f"{k}: {v}\n" for k, v in CFG.items())
                   + f"act_fn: relu\ndata_root: \"{data}/\"\nsplit: {{ 'test': \"4DMatch-F\" }}\naugment_noise: False\n"
                   "exp_dir: !join [vis, run]\nfolder: pyramid_level\n")
    res, stdout = _drive("eval_nolearned.py", tmp_path, tmp_path / "out.npz", argv=["--config", str(cfg)], emu=True)
    mods = [str(m) for m in res["modules"]]
    assert all("deformationpyramid_b200" in m for m in mods[:3]), mods          # the hot path is the mirror ...
    assert mods[3].startswith(REF), mods                                          # ... model.geometry stays the reference's
    assert "registration" in [str(k) for k in res["timer_keys"]]
    # replay: the script seeds once (setup_seed(0), eval_nolearned.py:22) and registers the pairs of both splits in
    # glob order, each pair drawing its weights and permutations from the global generator (registration.py:133-157)
    torch.manual_seed(0)
    for split in ("4DMatch-F", "4DLoMatch-F"):
        sums, cnt = {}, 0
        for path in [str(e) for e in res["entries_" + split]]:
            with np.load(path) as e:
                src, tgt = torch.from_numpy(e["s_pc"]), torch.from_numpy(e["t_pc"])
                rot, trans, flow, corr = e["rot"], e["trans"], e["s2t_flow"], e["correspondences"]
            ref = O.optimize_pair(O.NDPConfig(**{k: v for k, v in CFG.items()}), src, tgt)
            gt = torch.from_numpy(((rot @ (src.numpy() + flow).T + trans).T - src.numpy()).astype(np.float32))
            ov = np.zeros(len(src), bool); ov[corr[:, 0]] = True
            m = O.compute_flow_metrics(ref.warped - src, gt, overlap=torch.from_numpy(ov))
            for k, v in m.items():
                sums[k] = sums.get(k, 0.0) + v
            cnt += 1
        log = str(res["log_" + split][0])
        got = {k: float(v) for k, v in re.findall(r"([a-z]+-[A-Za-z]+): ([-0-9.naninf]+)", log)}
        assert set(got) == set(sums), (got, sums)
        for k in sums:
            want = sums[k] / cnt
            assert abs(got[k] - want) <= 2e-3 * max(1.0, abs(want)) + 6e-4, (split, k, got[k], want)   # the log prints 3 decimals


def _icosphere_ply(path, scale, bump, n_lat=10, n_lon=14):
    from deformationpyramid_b200.shape_transfer import write_ply_ascii
    v, f = [], []
    for i in range(n_lat + 1):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            r = scale * (1.0 + bump * np.sin(3 * th) * np.cos(2 * ph))
            v.append((r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), 1.3 * r * np.cos(th)))
    for i in range(n_lat):
        for j in range(n_lon):
            a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
            f.append((a, b, a + n_lon)); f.append((b, b + n_lon, a + n_lon))
    write_ply_ascii(path, np.asarray(v, np.float32), np.asarray(f, np.int64))


def test_shape_transfer_runs_unmodified_against_the_mirror(tmp_path):
    s_ply, t_ply = str(tmp_path / "s.ply"), str(tmp_path / "t.ply")
    _icosphere_ply(s_ply, 0.5, 0.10)
    _icosphere_ply(t_ply, 0.62, 0.18)
    over = dict(samples=128, m=2, iters=3)       # shape_transfer.py:27-49 hard-codes 6000 / 9 / 500: shrunk through the EasyDict stand-in
    res, _ = _drive("shape_transfer.py", tmp_path, tmp_path / "out.npz", argv=["-s", s_ply, "-t", t_ply], emu=True,
                    overrides=over, seed=3)
    assert int(res["drawn"][0]) == 3             # the three o3d.visualization.draw_geometries calls went to the stand-in
    src = torch.from_numpy(res["src_pcd"] + res["src_mean"])          # the sampled clouds before centring (:104-107)
    tgt = torch.from_numpy(res["tgt_pcd"] + res["tgt_mean"])
    cfg = O.NDPConfig(iters=3, samples=128, m=2, motion_type="Sim3", rotation_format="euler", max_break_count=15,
                      break_threshold_ratio=0.001)
    torch.manual_seed(0)                         # setup_seed(0) at import, then the pyramid is built (:22, 92-99)
    specs = O.make_specs(3, 128, -8, 2, "euler", motion="Sim3")
    init = [O.init_params(s) for s in specs]
    ref = O.optimize_pair(cfg, src, tgt, init=init, src_perm=torch.arange(128), tgt_perm=torch.arange(128))
    want, _ = O.pyramid_warp(specs, ref.params, torch.from_numpy(res["mesh_vert"]))     # mesh_vert is already centred (:163-165)
    err = np.abs(res["warped_vert"] - want.numpy()).max() / np.abs(want.numpy()).max()
    assert err < 1e-3, err
