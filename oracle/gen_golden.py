"""ORACLE - TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by importing and running the UNMODIFIED reference modules from
/root/reference (read-only) in the build container:
    model/nets.py, model/rigid_body.py         -- as they are (import fine under torch 2.11)
    model/loss.py, model/registration.py       -- as they are, with oracle/ref_stubs.py supplying
                                                  the absent imports (pytorch3d K=1 kNN = the C
                                                  loop of oracle/knn_oracle.c, fma rounding)
    torch.optim.Adam                           -- upstream torch, the class the reference calls
/root/reference does not exist on the GPU box, so the vectors travel as fixtures.  Run:
    python -m oracle.gen_golden            (from the repo root; takes ~1 minute)
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_stubs, ndp_oracle  # noqa: E402
from deformationpyramid_b200.synthetic import make_pair  # noqa: E402

ref_stubs.install("/root/reference")
import model.nets as ref_nets  # noqa: E402   (the reference's, from /root/reference)
import model.loss as ref_loss  # noqa: E402
import model.registration as ref_reg  # noqa: E402

assert ref_nets.__file__.startswith("/root/reference/"), ref_nets.__file__
GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def flat(params):
    return torch.cat([p.detach().reshape(-1) for p in params]).numpy().astype(np.float32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------
# G1: NDPLayer forward + autograd gradients for the option space of nets.py:10-62
# ------------------------------------------------------------------------------------
LAYER_VARIANTS = [
    # (motion, rotation_format, nonrigidity_est, level_m, depth)
    # depth 3 (NDP.yaml / shape_transfer.py) for the two shipped option sets, depth 2 for the
    # rest of the option space to keep the fixture small
    ("SE3", "axis_angle", False, 1, 3), ("SE3", "axis_angle", False, 9, 2), ("SE3", "euler", False, 3, 2),
    ("SE3", "quaternion", False, 4, 2), ("SE3", "6D", False, 5, 2),
    ("Sim3", "euler", False, 2, 3), ("Sim3", "axis_angle", False, 6, 2), ("Sim3", "quaternion", False, 7, 2),
    ("Sim3", "6D", False, 8, 2), ("sflow", "axis_angle", False, 4, 2),
    ("SE3", "axis_angle", True, 5, 2), ("Sim3", "euler", True, 3, 2), ("sflow", "axis_angle", True, 2, 2),
]


def gen_layers():
    out = {}
    meta = []
    for vi, (motion, fmt, nr, m, depth) in enumerate(LAYER_VARIANTS):
        seed = 100 + vi
        torch.manual_seed(seed)
        layer = ref_nets.NDPLayer(depth, 128, -8, m, fmt, nonrigidity_est=nr, motion=motion)
        g = torch.Generator().manual_seed(7000 + vi)
        n = 160
        # metre-scale points; higher levels see sub-metre structure
        x = (torch.rand(n, 3, generator=g) - 0.5) * 2.0
        # perturb the weights away from the near-identity init so that every branch matters
        with torch.no_grad():
            for p in layer.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=g))
            # rotation / scale heads: make angles O(0.3) so that rotation formulas are exercised
            layer.trn_branch.bias.add_(torch.randn(3, generator=g) * 50.0)
            if hasattr(layer, "rot_brach"):
                layer.rot_brach.bias.add_(torch.randn(layer.rot_brach.bias.shape, generator=g) * 300.0)
            if hasattr(layer, "s_branch"):
                layer.s_branch.bias.add_(100.0)
            if nr:
                layer.nr_branch.bias.add_(300.0)
        x.requires_grad_(True)
        y, nu = layer(x)
        gy = torch.randn(n, 3, generator=g)
        obj = (y * gy).sum()
        gnu = None
        if nu is not None:
            gnu = torch.randn(n, generator=g)
            obj = obj + (nu * gnu).sum()
        grads = torch.autograd.grad(obj, [x] + list(layer.parameters()))
        k = f"v{vi}"
        out[f"{k}_x"] = x.detach().numpy()
        out[f"{k}_params"] = flat(layer.parameters())
        out[f"{k}_y"] = y.detach().numpy()
        if nu is not None:
            out[f"{k}_nu"] = nu.detach().numpy()
            out[f"{k}_gnu"] = gnu.numpy()
        out[f"{k}_gy"] = gy.numpy()
        out[f"{k}_gx"] = grads[0].numpy()
        out[f"{k}_gparams"] = flat(grads[1:])
        meta.append(f"{motion},{fmt},{int(nr)},{m},{seed},{depth}")
        # fresh-init hash: pins that the repo's own constructors consume the RNG identically
        torch.manual_seed(seed)
        fresh = ref_nets.NDPLayer(depth, 128, -8, m, fmt, nonrigidity_est=nr, motion=motion)
        out[f"{k}_init_sha"] = np.array(sha(flat(fresh.parameters())))
        out[f"{k}_names"] = np.array([n_ for n_, _ in fresh.named_parameters()])
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(GOLD, "layers.npz"), **out)
    print("layers.npz", len(LAYER_VARIANTS), "variants")


# ------------------------------------------------------------------------------------
# G2: compute_truncated_chamfer_distance (loss.py:94-258) incl. the adversarial cases
# ------------------------------------------------------------------------------------
def gen_chamfer():
    out, meta = {}, []
    g = torch.Generator().manual_seed(42)

    def case(name, x, y, trunc):
        x = x.clone().requires_grad_(True)
        ref_stubs.KNN_MODE = 0
        loss = ref_loss.compute_truncated_chamfer_distance(x[None], y[None], trunc=trunc)
        gx, = torch.autograd.grad(loss, x)
        d2x, ix = ndp_oracle.knn1(x, y, 0)
        d2y, iy = ndp_oracle.knn1(y, x, 0)
        out[f"{name}_x"] = x.detach().numpy()
        out[f"{name}_y"] = y.numpy()
        out[f"{name}_trunc"] = np.float32(trunc)
        out[f"{name}_loss"] = loss.detach().numpy()
        out[f"{name}_gx"] = gx.numpy()
        out[f"{name}_d2x"], out[f"{name}_ix"] = d2x.numpy(), ix.numpy()
        out[f"{name}_d2y"], out[f"{name}_iy"] = d2y.numpy(), iy.numpy()
        meta.append(name)

    case("rand_small", torch.randn(37, 3, generator=g), torch.randn(53, 3, generator=g), 1e9)
    case("rand_ragged", torch.randn(1000, 3, generator=g) * 0.5, torch.randn(777, 3, generator=g) * 0.5, 1e9)
    # truncation active (LNDP.yaml trunc_cd = 0.25 on SQUARED distances)
    case("trunc", torch.randn(300, 3, generator=g), torch.randn(260, 3, generator=g), 0.25)
    # exact ties: every target duplicated (lowest index must win) and lattice points
    y = torch.randn(100, 3, generator=g)
    case("dup_targets", torch.randn(150, 3, generator=g), torch.cat([y, y, y]), 1e9)
    lat = torch.stack(torch.meshgrid(*[torch.arange(6.0)] * 3, indexing="ij"), -1).reshape(-1, 3)
    case("lattice", lat[torch.randperm(216, generator=g)][:128] + 0.5, lat, 1e9)
    src, tgt = make_pair(3, 2048, 2048)
    case("synth2048", src, tgt, 1e9)
    case("one_target", torch.randn(9, 3, generator=g), torch.randn(1, 3, generator=g), 1e9)
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(GOLD, "chamfer.npz"), **out)
    print("chamfer.npz", meta)


# ------------------------------------------------------------------------------------
# G3: teacher-forcing states along an unmodified Registration.register() trajectory
# ------------------------------------------------------------------------------------
class _Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def ndp_config(**kw):
    c = _Cfg(gpu_mode=False, deformation_model="NDP", use_ldmk=False, use_depth=False, iters=500,
             lr=0.01, max_break_count=15, break_threshold_ratio=0.001, w_reg=0.0, samples=2000,
             m=9, k0=-8, depth=3, width=128, act_fn="relu", motion_type="SE3",
             rotation_format="axis_angle", device=torch.device("cpu"))
    c.update(kw)
    return c


class _RecAdam(torch.optim.Adam):
    """torch.optim.Adam (the class registration.py:176 instantiates) that records, around every
    step, the state needed to teacher-force one iteration of a replacement."""
    log = []
    keep = None   # callable (level, it) -> bool
    level = -1
    steps = []    # Adam steps taken per level

    def __init__(self, params, **kw):
        super().__init__(params, **kw)
        _RecAdam.level += 1
        _RecAdam.steps.append(0)
        self._it = 0

    def _state(self):
        ps = [p for gr in self.param_groups for p in gr["params"]]
        m = [self.state[p]["exp_avg"] if p in self.state and "exp_avg" in self.state[p] else torch.zeros_like(p) for p in ps]
        v = [self.state[p]["exp_avg_sq"] if p in self.state and "exp_avg_sq" in self.state[p] else torch.zeros_like(p) for p in ps]
        st = [int(self.state[p]["step"]) if p in self.state and "step" in self.state[p] else 0 for p in ps]
        return ps, m, v, st[0] if st else 0

    def step(self, closure=None):
        rec = _RecAdam.keep is not None and _RecAdam.keep(_RecAdam.level, self._it)
        if rec:
            ps, m, v, st = self._state()
            r = dict(level=_RecAdam.level, it=self._it, step_before=st, params_before=flat(ps),
                     m_before=flat(m), v_before=flat(v), grads=flat([p.grad for p in ps]))
        res = super().step(closure)
        _RecAdam.steps[_RecAdam.level] += 1
        if rec:
            ps, m, v, st = self._state()
            r.update(params_after=flat(ps), m_after=flat(m), v_after=flat(v))
            _RecAdam.log.append(r)
        self._it += 1
        return res


def run_reference_register(cfg, src, tgt, seed, keep=None, landmarks=None):
    """Unmodified Registration.register() with instrumentation wrapped AROUND the reference's
    own calls (the recording Adam subclass and a recording wrapper of the Chamfer function)."""
    _RecAdam.log, _RecAdam.keep, _RecAdam.level, _RecAdam.steps = [], keep, -1, []
    cd_log = []
    orig_cd = ref_loss.compute_truncated_chamfer_distance

    def rec_cd(x, y, **kw):
        loss = orig_cd(x, y, **kw)
        cd_log.append((x.detach()[0].clone(), y.detach()[0].clone(), float(loss)))
        return loss

    ref_reg.optim.Adam = _RecAdam
    ref_reg.compute_truncated_chamfer_distance = rec_cd
    try:
        torch.manual_seed(seed)
        reg = ref_reg.Registration(cfg)
        reg.load_pcds(src.numpy(), tgt.numpy(), landmarks=landmarks)
        warped, _, _ = reg.register()
    finally:
        ref_reg.optim.Adam = torch.optim.Adam
        ref_reg.compute_truncated_chamfer_distance = orig_cd
    return warped.detach(), cd_log, list(_RecAdam.log)


def gen_trajectory():
    """Config-2 shape scaled to fixture size: one level, early stop off, N = M = 512 samples."""
    src, tgt = make_pair(11, 512, 512)
    cfg = ndp_config(m=1, iters=60, samples=512, max_break_count=10 ** 9)
    keep_its = (0, 1, 2, 30, 31)
    warped, cd_log, adam_log = run_reference_register(
        cfg, src, tgt, seed=5, keep=lambda lv, it: it in keep_its)
    out = dict(src=src.numpy(), tgt=tgt.numpy(), seed=np.int64(5), warped=warped.numpy(),
               keep_its=np.array(keep_its), losses=np.array([c[2] for c in cd_log], np.float32),
               t_sample=cd_log[0][1].numpy())
    for r in adam_log:
        k = f"it{r['it']}"
        for name in ("params_before", "m_before", "v_before", "grads", "params_after", "m_after",
                     "v_after"):
            out[f"{k}_{name}"] = r[name]
        out[f"{k}_step_before"] = np.int64(r["step_before"])
        out[f"{k}_x_warped"] = cd_log[r["it"]][0].numpy()
    # the level input (s_sample) is recovered as in registration.py:150-159 by the test
    np.savez_compressed(os.path.join(GOLD, "trajectory.npz"), **out)
    print("trajectory.npz", [r["it"] for r in adam_log], "final loss", cd_log[-1][2])


# ------------------------------------------------------------------------------------
# G4: whole-pair runs of the unmodified Registration.register()
# ------------------------------------------------------------------------------------
def gen_pairs():
    out, meta = {}, []

    def case(name, cfg, p, n, m, seed):
        src, tgt = make_pair(p, n, m)
        warped, cd_log, _ = run_reference_register(cfg, src, tgt, seed)
        out[f"{name}_pair"] = np.array([p, n, m, seed])
        out[f"{name}_warped"] = warped.numpy()
        out[f"{name}_losses"] = np.array([c[2] for c in cd_log], np.float32)
        out[f"{name}_cfg"] = np.array([f"{k}={v}" for k, v in cfg.items() if k != "device"])
        meta.append(name)
        print(name, "iterations", len(cd_log), "final loss", cd_log[-1][2])

    # as-configured early stop, full 9-level pyramid, small clouds / few samples
    case("ndp9", ndp_config(samples=256, iters=40), 21, 700, 650, 1)
    # fixed-iteration mode (early stop disabled), 3 levels
    case("fixed3", ndp_config(m=3, samples=300, iters=25, max_break_count=10 ** 9), 22, 300, 420, 2)
    # shape_transfer.py's model options (Sim3 + euler), through the same driver
    case("sim3euler", ndp_config(m=4, samples=256, iters=20, motion_type="Sim3",
                                 rotation_format="euler"), 23, 400, 400, 3)
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(GOLD, "pairs.npz"), **out)


# ------------------------------------------------------------------------------------
# G5: evaluation metrics of the unmodified model/loss.py:382-403, 431-471
# ------------------------------------------------------------------------------------
def gen_metrics():
    out, names = {}, []
    g = torch.Generator().manual_seed(9)
    for name, n, noise, frac in (("small", 40, 0.01, 0.5), ("mixed", 3000, 0.03, 0.7), ("bad", 500, 0.4, 0.2),
                                 ("allvis", 64, 0.02, 1.0)):
        gt = 0.1 * torch.randn(n, 3, generator=g)
        gt[: n // 8] *= 0.01                                   # tiny labels: the relative criteria decide
        flow = gt + noise * torch.randn(n, 3, generator=g) * torch.rand(n, 1, generator=g)
        overlap = torch.rand(n, generator=g) < frac
        m = ref_loss.compute_flow_metrics(flow, gt, overlap=overlap)
        m2 = ref_loss.compute_flow_metrics(flow, gt)
        s = ref_loss.scene_flow_metrics(flow, gt)
        out[f"{name}_flow"], out[f"{name}_gt"], out[f"{name}_overlap"] = flow.numpy(), gt.numpy(), overlap.numpy()
        out[f"{name}_keys"] = np.array(list(m.keys()))
        out[f"{name}_vals"] = np.array([m[k] for k in m], np.float64)
        out[f"{name}_keys_noov"] = np.array(list(m2.keys()))
        out[f"{name}_vals_noov"] = np.array([m2[k] for k in m2], np.float64)
        out[f"{name}_scene"] = np.array(s, np.float64)
        names.append(name)
    out["meta"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **out)
    print("metrics.npz", names)


# ------------------------------------------------------------------------------------
# G6: the other branches of the unmodified Registration.register(): landmarks (LNDP, registration.py:187-203,
#     with and without the truncated Chamfer term) and the nonrigidity regulariser (:216-220)
# ------------------------------------------------------------------------------------
def gen_branches():
    out, meta = {}, []

    def case(name, cfg, p, n, m, seed, n_ldmk):
        src, tgt = make_pair(p, n, m)
        landmarks = None
        if n_ldmk:
            g = torch.Generator().manual_seed(100 + p)
            pick = torch.randperm(n, generator=g)[:n_ldmk]
            ls = src[pick].clone()
            # plausible correspondences: the source landmarks pushed through the pair generator's own smooth warp
            # (a rotation about z, a shift and a sinusoidal displacement), plus a little noise
            ang = 0.17
            R = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
            lt = ls @ R.T + torch.tensor([0.03, -0.02, 0.01]) + 0.04 * torch.sin(3.0 * ls[:, [1, 2, 0]])
            lt = lt + 0.002 * torch.randn(n_ldmk, 3, generator=g)
            landmarks = (ls, lt)
            out[f"{name}_ldmk_s"], out[f"{name}_ldmk_t"] = ls.numpy(), lt.numpy()
        warped, cd_log, _ = run_reference_register(cfg, src, tgt, seed, landmarks=landmarks)
        out[f"{name}_pair"] = np.array([p, n, m, seed, n_ldmk])
        out[f"{name}_warped"] = warped.numpy()
        out[f"{name}_steps"] = np.array(_RecAdam.steps, np.int64)
        out[f"{name}_cd"] = np.array([c[2] for c in cd_log], np.float32)
        out[f"{name}_cfg"] = np.array([f"{k}={v}" for k, v in cfg.items() if k != "device"])
        meta.append(name)
        print(name, "adam steps per level", _RecAdam.steps, "chamfer calls", len(cd_log))

    # config/LNDP.yaml's shape (w_cd = 0: landmark loss only), 3 levels
    case("ldmk", ndp_config(m=3, samples=256, iters=25, w_cd=0.0, trunc_cd=0.25), 41, 500, 480, 4, 60)
    # landmark + truncated Chamfer (registration.py:189-197; trunc on SQUARED distances)
    case("ldmk_cd", ndp_config(m=3, samples=256, iters=20, w_cd=0.1, trunc_cd=0.0025), 42, 500, 480, 5, 60)
    # nonrigidity regulariser (registration.py:216-220): nr_branch on every level but the first
    case("wreg", ndp_config(m=3, samples=256, iters=20, w_reg=0.2, max_break_count=10 ** 9), 43, 400, 380, 6, 0)
    out["meta"] = np.array(meta)
    np.savez_compressed(os.path.join(GOLD, "branches.npz"), **out)


# ------------------------------------------------------------------------------------
# G7: BASELINE.json config 2 AS WRITTEN (2048-pt pair, one level, 200 Adam iterations, early stop off):
#     the reference's loss curve, sample permutations and final warped cloud (the per-step states are
#     regenerated by the tests from the oracle, which this curve pins over all 200 steps)
# ------------------------------------------------------------------------------------
def gen_config2():
    src, tgt = make_pair(12, 2048, 2048)
    cfg = ndp_config(m=1, iters=200, samples=2048, max_break_count=10 ** 9)
    warped, cd_log, _ = run_reference_register(cfg, src, tgt, seed=7)
    out = dict(pair=np.array([12, 2048, 2048, 7]), losses=np.array([c[2] for c in cd_log], np.float64),
               warped=warped.numpy(), x_last=cd_log[-1][0].numpy(), t_sample=cd_log[0][1].numpy(),
               x_it100=cd_log[100][0].numpy())
    np.savez_compressed(os.path.join(GOLD, "config2.npz"), **out)
    print("config2.npz", len(cd_log), "iterations, loss", cd_log[0][2], "->", cd_log[-1][2])


# ------------------------------------------------------------------------------------
# G8: the reference's own fixtures for BASELINE.json config 1 (sim3_demo/*.ply), gzip-compressed byte for byte
# ------------------------------------------------------------------------------------
def gen_meshes():
    import gzip
    for name in ("AlienSoldier", "Ortiz"):
        with open(f"/root/reference/sim3_demo/{name}.ply", "rb") as f:
            raw = f.read()
        with open(os.path.join(GOLD, f"{name}.ply.gz"), "wb") as f, gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as z:
            z.write(raw)
        print(name, len(raw), "bytes, sha256", hashlib.sha256(raw).hexdigest()[:16])


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for fn in (gen_layers, gen_chamfer, gen_trajectory, gen_pairs, gen_metrics, gen_branches, gen_config2, gen_meshes):
        if not only or fn.__name__ in only:
            fn()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")
