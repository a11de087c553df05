"""Headless shape transfer: the pre/post steps of the reference's shape_transfer.py
without open3d, around the same optimisation loop.

  read_ply_ascii / write_ply_ascii   -- the ASCII PLY meshes of sim3_demo/ (x y z nx ny nz s t + face lists);
                                        replaces o3d.io.read_triangle_mesh (shape_transfer.py:69,78)
  sample_points_uniformly            -- area-weighted uniform surface sampling, replaces
                                        mesh.sample_points_uniformly(number_of_points=6000) (:71,80)
  shape_transfer                     -- shape_transfer.py:92-165: Sim3 + euler pyramid, Chamfer-only objective on
                                        all sampled points, then the 9-level inference warp of every mesh vertex

    python -m deformationpyramid_b200.shape_transfer -s AlienSoldier.ply -t Ortiz.ply [-o out.ply]

The loop itself runs in the fused CUDA solver (same kernels as Registration.register); the script's
verbatim Python control flow over the autograd ops is exercised by tests/test_gpu_api.py.
"""
from __future__ import annotations

import argparse
from typing import Optional, Tuple

import numpy as np
import torch

from .config import AttrDict

# shape_transfer.py:27-49 (verbatim values)
SHAPE_TRANSFER_CONFIG = dict(gpu_mode=True, iters=500, lr=0.01, max_break_count=15, break_threshold_ratio=0.001,
                             samples=6000, motion_type="Sim3", rotation_format="euler", m=9, k0=-8, depth=3,
                             width=128, act_fn="relu", w_reg=0, w_ldmk=0, w_cd=0.1, deformation_model="NDP")


def read_ply_ascii(path: str) -> Tuple[np.ndarray, np.ndarray, list]:
    """-> (vertices [V,3] float32, faces [F,3] int64 (polygons fan-triangulated), header property names)."""
    import gzip
    opener = (lambda q: gzip.open(q, "rt")) if str(path).endswith(".gz") else (lambda q: open(q, "r"))
    with opener(path) as f:
        if f.readline().strip() != "ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = f.readline().split()
        if fmt[:2] != ["format", "ascii"]:
            raise ValueError(f"{path}: only ASCII PLY is supported")
        nv = nf = 0
        props, cur = [], None
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "element":
                cur = tok[1]
                if cur == "vertex":
                    nv = int(tok[2])
                elif cur == "face":
                    nf = int(tok[2])
            elif tok[0] == "property" and cur == "vertex":
                props.append(tok[-1])
            elif tok[0] == "end_header":
                break
        ix, iy, iz = props.index("x"), props.index("y"), props.index("z")
        verts = np.empty((nv, 3), np.float32)
        for i in range(nv):
            t = f.readline().split()
            verts[i] = (float(t[ix]), float(t[iy]), float(t[iz]))
        faces = []
        for _ in range(nf):
            t = f.readline().split()
            k = int(t[0])
            idx = [int(v) for v in t[1:1 + k]]
            for j in range(1, k - 1):
                faces.append((idx[0], idx[j], idx[j + 1]))
    return verts, np.asarray(faces, np.int64).reshape(-1, 3), props


def write_ply_ascii(path: str, verts: np.ndarray, faces: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment written by deformationpyramid_b200\n")
        f.write(f"element vertex {len(verts)}\nproperty float x\nproperty float y\nproperty float z\n")
        f.write(f"element face {len(faces)}\nproperty list uchar uint vertex_indices\nend_header\n")
        for v in verts:
            f.write(f"{v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
        for t in faces:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")


def sample_points_uniformly(verts: np.ndarray, faces: np.ndarray, number_of_points: int,
                            rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Uniform samples on the surface: triangles drawn with probability proportional to their area,
    points uniform inside a triangle (sqrt trick)."""
    rng = rng if rng is not None else np.random.default_rng(0)
    a, b, c = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    if not np.isfinite(area).all() or area.sum() <= 0:
        raise ValueError("degenerate mesh")
    tri = rng.choice(len(faces), size=number_of_points, p=area / area.sum())
    r1, r2 = np.sqrt(rng.random(number_of_points)), rng.random(number_of_points)
    w = np.stack([1 - r1, r1 * (1 - r2), r1 * r2], 1).astype(np.float64)
    pts = w[:, 0:1] * a[tri] + w[:, 1:2] * b[tri] + w[:, 2:3] * c[tri]
    return pts.astype(np.float32)


def shape_transfer(src_pts: np.ndarray, tgt_pts: np.ndarray, src_verts: np.ndarray, device=0, seed: int = 0,
                   add_target_mean: bool = False, **overrides):
    """shape_transfer.py:92-165 for already sampled clouds.  Returns (warped vertices [V,3] numpy,
    Adam steps per level, last loss per level).  Like the reference (shape_transfer.py:161-165) the warped
    vertices stay in the CENTRED target frame; add_target_mean=True moves them onto the target mesh."""
    from .model.registration import Registration
    cfg = AttrDict(SHAPE_TRANSFER_CONFIG)
    for k, v in overrides.items():
        cfg[k] = v
    cfg.device = device
    # the script optimises on ALL sampled points (s_sample = src_pcd, :112-113) and finally warps every
    # mesh vertex through the pyramid (:161-165).  Registration.register() warps the cloud it was given,
    # so the vertices ride along: the sampled points lead the cloud and are exactly the optimised subset.
    n_s = src_pts.shape[0]
    cfg.samples = max(n_s, tgt_pts.shape[0])
    cloud = np.concatenate([src_pts, src_verts]).astype(np.float32)
    torch.manual_seed(seed)
    reg = Registration(cfg)
    reg.load_pcds(cloud, tgt_pts.astype(np.float32))
    warped, _, _ = _register_leading_samples(reg, n_s, tgt_pts.shape[0], add_target_mean)
    return warped[n_s:].cpu().numpy(), reg.last_iters, reg.last_losses


def _register_leading_samples(reg, n_src_samples: int, n_tgt_samples: int, add_target_mean: bool = False):
    """The fused driver with the sub-sampling of registration.py:156-159 replaced by the identity on the
    leading n samples (shape_transfer.py optimises on all sampled points, :112-113) and the source centred
    on the mean of the SAMPLED points (shape_transfer.py:104-107, 163-164), not of the whole cloud.
    The native solver subtracts the mean of the cloud it is given, so the cloud is pre-centred here and one
    balancing point is appended that makes that mean zero; it is warped along and dropped.  The numbers of
    optimised points are passed to the solver explicitly (src_samples / tgt_samples of the C ABI), so clouds of
    different sizes are fine: config.samples is only the capacity."""
    NDP = reg._build_pyramid()
    src, tgt = reg.src_pcd.contiguous(), reg.tgt_pcd.contiguous()
    dev = src.device
    src_mean = src[:n_src_samples].mean(dim=0, keepdim=True)
    tgt_mean = tgt.mean(dim=0, keepdim=True)
    src_c = src - src_mean
    src_in = torch.cat([src_c, -src_c.sum(dim=0, keepdim=True)]).contiguous()
    tgt_in = (tgt - tgt_mean).contiguous()
    flat = NDP.flat_parameters()
    sp = torch.arange(n_src_samples, dtype=torch.int32, device=dev)
    tp = torch.arange(n_tgt_samples, dtype=torch.int32, device=dev)
    solver = reg._get_solver(1, src_in.shape[0], tgt_in.shape[0])
    warped, iters, losses = solver.register([src_in], [tgt_in], [flat], [sp], [tp], src_samples=[n_src_samples],
                                            tgt_samples=[n_tgt_samples])
    NDP.load_flat_parameters(flat)
    reg.NDP, reg.last_iters, reg.last_losses = NDP, iters[0], losses[0]
    out = warped[0][:-1]            # the native solver adds the mean of the target cloud it was given: zero here
    return (out + tgt_mean if add_target_mean else out), {}, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", type=str, required=True, help="Path to the src mesh.")
    ap.add_argument("-t", type=str, required=True, help="Path to the tgt mesh.")
    ap.add_argument("-o", type=str, default=None, help="Write the fitted mesh here (ASCII PLY).")
    ap.add_argument("--samples", type=int, default=SHAPE_TRANSFER_CONFIG["samples"])
    ap.add_argument("--on-target", action="store_true", help="add the target mean back (the reference leaves the "
                    "fitted vertices in the centred target frame, shape_transfer.py:161-165)")
    args = ap.parse_args()
    sv, sf, _ = read_ply_ascii(args.s)
    tv, tf, _ = read_ply_ascii(args.t)
    rng = np.random.default_rng(0)
    sp = sample_points_uniformly(sv, sf, args.samples, rng)
    tp = sample_points_uniformly(tv, tf, args.samples, rng)
    warped, iters, losses = shape_transfer(sp, tp, sv, add_target_mean=args.on_target)
    print("Adam steps per level:", [int(v) for v in iters], "last loss per level:", [round(float(v), 5) for v in losses])
    if args.o:
        write_ply_ascii(args.o, warped, sf)
        print("wrote", args.o)


if __name__ == "__main__":
    main()
