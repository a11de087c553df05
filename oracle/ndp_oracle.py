"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement (torch fp32 on CPU + the C nearest-neighbour loop of oracle/knn_oracle.c) of
the reference's per-pair Neural-Deformation-Pyramid hot path.  Nothing under
deformationpyramid_b200/ imports this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do, and there only as the checker / the CPU arm.

Parity status: PINNED.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the UNMODIFIED reference modules imported in
the build container by oracle/gen_golden.py (model/nets.py, model/rigid_body.py, model/loss.py
and model/registration.py from /root/reference, with pytorch3d's knn_points supplied by
oracle/knn_oracle.c) and committed as tests/golden/*.npz; tests/test_oracle_golden.py checks
every function below against them.

Each function cites the reference lines it follows.  The code is a restatement (functional,
explicit parameter dictionaries), not a copy.
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libndp_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_ndp_oracle_build",
                                                          os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            path = mod.build()
        lib = ctypes.CDLL(path)
        lib.ndp_oracle_knn1.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int]
        lib.ndp_oracle_knn1.restype = None
        lib.ndp_oracle_knn1_mode_disagreements.argtypes = [ctypes.c_void_p, ctypes.c_int64,
                                                           ctypes.c_void_p, ctypes.c_int64]
        lib.ndp_oracle_knn1_mode_disagreements.restype = ctypes.c_int64
        lib.ndp_oracle_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


# --------------------------------------------------------------------------------------
# Nearest neighbour (pytorch3d.ops.knn.knn_points, K=1)  [upstream, un-vendored]
# call sites: model/loss.py:177-178
# --------------------------------------------------------------------------------------
def knn1(p1: torch.Tensor, p2: torch.Tensor, mode: int = 0, threads: int = 1
         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """p1 [n,3], p2 [m,3] fp32 CPU -> (squared distance [n] fp32, index [n] int64).

    Ascending scan, strict '<' (lowest index wins ties), direct-difference squared L2.
    mode 0 = fma rounding (the parity contract), mode 1 = separately rounded mul/add."""
    p1 = p1.detach().to(torch.float32).contiguous().cpu()
    p2 = p2.detach().to(torch.float32).contiguous().cpu()
    n, m = p1.shape[0], p2.shape[0]
    if m < 1:
        raise ValueError("knn1: the searched cloud must not be empty")
    d2 = torch.empty(n, dtype=torch.float32)
    idx = torch.empty(n, dtype=torch.int64)
    if n:
        _lib().ndp_oracle_knn1(p1.data_ptr(), n, p2.data_ptr(), m, d2.data_ptr(),
                               idx.data_ptr(), int(mode), int(threads))
    return d2, idx


def knn1_mode_disagreements(p1: torch.Tensor, p2: torch.Tensor) -> int:
    p1 = p1.detach().to(torch.float32).contiguous().cpu()
    p2 = p2.detach().to(torch.float32).contiguous().cpu()
    return int(_lib().ndp_oracle_knn1_mode_disagreements(p1.data_ptr(), p1.shape[0],
                                                         p2.data_ptr(), p2.shape[0]))


def max_threads() -> int:
    return int(_lib().ndp_oracle_max_threads())


class _Knn1Fn(torch.autograd.Function):
    """Differentiable K=1 search with pytorch3d's backward [upstream knn.py:_knn_points.backward
    -> KNearestNeighborBackwardCpu]: with g = dL/d(dists),
        grad_p1[i]      += 2 g_i (p1_i - p2_idx(i))
        grad_p2[idx(i)] -= 2 g_i (p1_i - p2_idx(i))      (sequential scatter, ascending i)."""

    @staticmethod
    def forward(ctx, p1, p2, mode, threads):
        d2, idx = knn1(p1, p2, mode, threads)
        ctx.save_for_backward(p1, p2, idx)
        ctx.mark_non_differentiable(idx)
        return d2, idx

    @staticmethod
    def backward(ctx, g, _gidx):
        p1, p2, idx = ctx.saved_tensors
        diff = 2.0 * g[:, None] * (p1 - p2[idx])
        gp1 = diff
        gp2 = torch.zeros_like(p2).index_add_(0, idx, -diff)
        return gp1, gp2, None, None


def knn1_autograd(p1, p2, mode=0, threads=1):
    return _Knn1Fn.apply(p1, p2, mode, threads)


# --------------------------------------------------------------------------------------
# Truncated L1 Chamfer  (model/loss.py:94-258)
# --------------------------------------------------------------------------------------
def chamfer_truncated(x: torch.Tensor, y: torch.Tensor, trunc: float = 0.2,
                      mode: int = 0, threads: int = 1, return_nn: bool = False):
    """x [B,P1,3], y [B,P2,3] -> 0-dim loss (batch/point reduction "mean", no weights, no
    normals, homogeneous lengths: the only configuration the reference exercises).

    loss.py:177-181 both NN searches; :185-188 entries with SQUARED distance >= trunc are
    zeroed (and carry no gradient); :227-228 sqrt -> L1; :233-235 divide by the FULL lengths;
    :240-251 batch mean; :255 sum of both directions."""
    if x.ndim != 3 or y.ndim != 3:
        raise ValueError("Expected points to be of shape (N, P, D)")       # loss.py:39-40
    if y.shape[0] != x.shape[0] or y.shape[2] != x.shape[2]:
        raise ValueError("y does not have the correct shape.")              # loss.py:158-159
    B, P1, _ = x.shape
    P2 = y.shape[1]
    tot_x = x.new_zeros(())
    tot_y = x.new_zeros(())
    nn_out = []
    for b in range(B):
        d2x, ix = knn1_autograd(x[b], y[b], mode, threads)
        d2y, iy = knn1_autograd(y[b], x[b], mode, threads)
        cx = torch.where(d2x >= trunc, torch.zeros_like(d2x), d2x)
        cy = torch.where(d2y >= trunc, torch.zeros_like(d2y), d2y)
        tot_x = tot_x + torch.sqrt(cx).sum() / P1
        tot_y = tot_y + torch.sqrt(cy).sum() / P2
        nn_out.append((d2x.detach(), ix, d2y.detach(), iy))
    loss = tot_x / B + tot_y / B
    if return_nn:
        return loss, nn_out
    return loss


# --------------------------------------------------------------------------------------
# Rotation parameterisations  (model/rigid_body.py, model/nets.py:144-161)
# --------------------------------------------------------------------------------------
def rot_axis_angle(a: torch.Tensor) -> torch.Tensor:
    """nets.py:150-153 + rigid_body.py:113-119 (Rodrigues), a [N,3] -> R [N,3,3].
    theta = |a|, w = a/theta (NaN for a == 0, as in the reference)."""
    theta = torch.linalg.vector_norm(a, dim=-1, keepdim=True)
    w = a / theta
    z = torch.zeros_like(w[:, 0])
    K = torch.stack([z, -w[:, 2], w[:, 1],
                     w[:, 2], z, -w[:, 0],
                     -w[:, 1], w[:, 0], z], dim=-1).reshape(-1, 3, 3)     # rigid_body.py:89-95
    th = theta[..., None]
    eye = torch.eye(3, dtype=a.dtype)[None]
    return eye + torch.sin(th) * K + ((1.0 - torch.cos(th)) * K) @ K


def rot_euler(a: torch.Tensor) -> torch.Tensor:
    """nets.py:148-149 + rigid_body.py:19-56: R = Rx(a0) Ry(a1) Rz(a2)."""
    c, s = torch.cos(a), torch.sin(a)
    o, z = torch.ones_like(a[:, 0]), torch.zeros_like(a[:, 0])
    Rx = torch.stack([o, z, z, z, c[:, 0], -s[:, 0], z, s[:, 0], c[:, 0]], -1).reshape(-1, 3, 3)
    Ry = torch.stack([c[:, 1], z, s[:, 1], z, o, z, -s[:, 1], z, c[:, 1]], -1).reshape(-1, 3, 3)
    Rz = torch.stack([c[:, 2], -s[:, 2], z, s[:, 2], c[:, 2], z, z, z, o], -1).reshape(-1, 3, 3)
    return (Rx @ Ry) @ Rz


def rot_quaternion(a: torch.Tensor) -> torch.Tensor:
    """nets.py:154-157 + rigid_body.py:58-85.  q = a / copysign(|a|, a0); R(q) with
    two_s = 2/(q.q) (kept as a differentiable function of q, as in the reference)."""
    nrm = torch.sqrt((a * a).sum(1))
    den = torch.where(a[:, 0] < 0, -nrm, nrm)                              # rigid_body.py:58-60
    q = a / den[:, None]
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    R = torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)],
                    -1)
    return R.reshape(-1, 3, 3)


def rot_6d(a: torch.Tensor) -> torch.Tensor:
    """nets.py:158-159 + rigid_body.py:5-16 (Gram-Schmidt, rows b1,b2,b3; normalize eps 1e-12)."""
    a1, a2 = a[:, :3], a[:, 3:]
    b1 = a1 / torch.clamp(torch.linalg.vector_norm(a1, dim=-1, keepdim=True), min=1e-12)
    u2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = u2 / torch.clamp(torch.linalg.vector_norm(u2, dim=-1, keepdim=True), min=1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


ROT_DIM = {"axis_angle": 3, "euler": 3, "quaternion": 4, "6D": 6}
_ROT_FN = {"axis_angle": rot_axis_angle, "euler": rot_euler, "quaternion": rot_quaternion,
           "6D": rot_6d}


# --------------------------------------------------------------------------------------
# One pyramid level  (model/nets.py:66-183, 295-304)
# --------------------------------------------------------------------------------------
@dataclass
class LayerSpec:
    depth: int = 3
    width: int = 128
    k0: int = -8
    m: int = 1                      # 1-based level number, nets.py:25 passes i+1
    rotation_format: str = "axis_angle"
    nonrigidity_est: bool = False
    motion: str = "SE3"
    mlp_scale: float = 0.001        # nets.py:107

    @property
    def freq(self) -> float:        # nets.py:168 (no pi factor)
        return float(2.0 ** (self.m + self.k0))


def param_layout(spec: LayerSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in nn.Module.parameters() order of the reference NDPLayer
    (nets.py:75-103: input, mlp, rot_brach, s_branch, trn_branch, nr_branch)."""
    W = spec.width
    out = [("input.0.weight", (W, 6)), ("input.0.bias", (W,))]
    for l in range(spec.depth - 1):
        out += [(f"mlp.pts_linears.{l}.weight", (W, W)), (f"mlp.pts_linears.{l}.bias", (W,))]
    if spec.motion in ("SE3", "Sim3"):
        R = ROT_DIM[spec.rotation_format]
        out += [("rot_brach.weight", (R, W)), ("rot_brach.bias", (R,))]
        if spec.motion == "Sim3":
            out += [("s_branch.weight", (1, W)), ("s_branch.bias", (1,))]
    out += [("trn_branch.weight", (3, W)), ("trn_branch.bias", (3,))]
    if spec.nonrigidity_est:
        out += [("nr_branch.weight", (1, W)), ("nr_branch.bias", (1,))]
    return out


def param_count(spec: LayerSpec) -> int:
    return sum(int(np.prod(s)) for _, s in param_layout(spec))


def init_params(spec: LayerSpec) -> Dict[str, torch.Tensor]:
    """Fresh weights drawn from the global torch CPU generator in the reference's order:
    nn.Linear default init per sub-module in construction order (nets.py:75-103), then Xavier
    uniform over every parameter with dim > 1 in parameters() order (nets.py:180-183)."""
    mods = {}
    for name, shape in param_layout(spec):
        if name.endswith(".weight"):
            mods[name[:-7]] = torch.nn.Linear(shape[1], shape[0])
    P = {}
    for name, shape in param_layout(spec):
        base, leaf = name.rsplit(".", 1)
        P[name] = getattr(mods[base], leaf).detach().clone()
    for name, _ in param_layout(spec):
        if P[name].dim() > 1:
            torch.nn.init.xavier_uniform_(P[name])
    return P


def flatten_params(spec: LayerSpec, P: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([P[n].reshape(-1) for n, _ in param_layout(spec)])


def unflatten_params(spec: LayerSpec, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
    P, off = {}, 0
    for n, s in param_layout(spec):
        k = int(np.prod(s))
        P[n] = flat[off:off + k].reshape(s)
        off += k
    return P


def posenc(spec: LayerSpec, x: torch.Tensor) -> torch.Tensor:
    """nets.py:164-177: [sin fx, cos fx, sin fy, cos fy, sin fz, cos fz], f = 2**(m+k0)."""
    fx = x * spec.freq
    return torch.stack([torch.sin(fx), torch.cos(fx)], dim=-1).reshape(x.shape[0], 6)


def layer_hidden(spec: LayerSpec, P: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """nets.py:113-115 + 295-304: input Linear+ReLU, then depth-1 Linear+ReLU."""
    h = torch.relu(posenc(spec, x) @ P["input.0.weight"].T + P["input.0.bias"])
    for l in range(spec.depth - 1):
        h = torch.relu(h @ P[f"mlp.pts_linears.{l}.weight"].T + P[f"mlp.pts_linears.{l}.bias"])
    return h


def layer_forward(spec: LayerSpec, P: Dict[str, torch.Tensor], x: torch.Tensor
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """NDPLayer.forward, nets.py:111-140.  x [N,3] -> (x' [N,3], nonrigidity [N] | None)."""
    mu = spec.mlp_scale
    h = layer_hidden(spec, P, x)
    t = mu * (h @ P["trn_branch.weight"].T + P["trn_branch.bias"])                 # :117
    if spec.motion in ("SE3", "Sim3"):
        a = mu * (h @ P["rot_brach.weight"].T + P["rot_brach.bias"])               # :146
        R = _ROT_FN[spec.rotation_format](a)
        Rx = (R @ x[:, :, None])[:, :, 0]
        if spec.motion == "SE3":
            y = Rx + t                                                             # :121
        else:
            s = mu * (h @ P["s_branch.weight"].T + P["s_branch.bias"]) + 1.0       # :125
            y = s * Rx + t                                                         # :126
    else:
        y = x + t                                                                  # :129
    nr = None
    if spec.nonrigidity_est:
        nu = torch.sigmoid(mu * (h @ P["nr_branch.weight"].T + P["nr_branch.bias"]))  # :133
        y = x + nu * (y - x)                                                       # :134
        nr = nu[:, 0]
    return y, nr


def pyramid_warp(specs: Sequence[LayerSpec], params: Sequence[Dict[str, torch.Tensor]],
                 x: torch.Tensor, max_level: Optional[int] = None, min_level: int = 0):
    """Deformation_Pyramid.warp, nets.py:36-48."""
    if max_level is None:
        max_level = len(specs) - 1
    assert max_level < len(specs), "more level than defined"
    data = {}
    for i in range(min_level, max_level + 1):
        x, nr = layer_forward(specs[i], params[i], x)
        data[i] = (x, nr)
    return x, data


def make_specs(depth, width, k0, m, rotation_format, nonrigidity_est=False, motion="SE3"
               ) -> List[LayerSpec]:
    """Deformation_Pyramid.__init__, nets.py:12-33 (nonrigidity only above level 0, :27)."""
    assert motion in ["Sim3", "SE3", "sflow"]
    return [LayerSpec(depth, width, k0, i + 1, rotation_format,
                      bool(nonrigidity_est) and i != 0, motion) for i in range(m)]


# --------------------------------------------------------------------------------------
# Early stop  (model/registration.py:225-232; identical in shape_transfer.py:142-149)
# --------------------------------------------------------------------------------------
@dataclass
class EarlyStop:
    max_break_count: int
    break_threshold_ratio: float
    break_counter: int = 0          # cumulative per level, never reset (registration.py:179)
    loss_prev: float = 1e6          # registration.py:180

    def should_stop(self, loss: float) -> bool:
        """Python-float (fp64) tests on the fp32 loss value, evaluated after forward+loss and
        before backward/step."""
        if loss < 1e-4:
            return True
        if abs(self.loss_prev - loss) < self.loss_prev * self.break_threshold_ratio:
            self.break_counter += 1
        if self.break_counter >= self.max_break_count:
            return True
        self.loss_prev = loss
        return False


# --------------------------------------------------------------------------------------
# Per-pair driver  (model/registration.py:126-262)
# --------------------------------------------------------------------------------------
@dataclass
class NDPConfig:
    """The keys of config/NDP.yaml that registration.py reads (:128-140,158,176,184,216)."""
    iters: int = 500
    lr: float = 0.01
    max_break_count: int = 15
    break_threshold_ratio: float = 0.001
    w_reg: float = 0.0
    samples: int = 2000
    m: int = 9
    k0: int = -8
    depth: int = 3
    width: int = 128
    motion_type: str = "SE3"
    rotation_format: str = "axis_angle"
    w_cd: float = 0.0
    trunc_cd: float = 0.25


@dataclass
class PairResult:
    warped: torch.Tensor
    iters_per_level: List[int]
    loss_per_level: List[float]
    loss_curve: List[List[float]] = field(default_factory=list)


def optimize_pair(cfg: NDPConfig, src_pcd: torch.Tensor, tgt_pcd: torch.Tensor,
                  init: Optional[Sequence[Dict[str, torch.Tensor]]] = None,
                  src_perm: Optional[torch.Tensor] = None, tgt_perm: Optional[torch.Tensor] = None,
                  landmarks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                  knn_mode: int = 0, knn_threads: int = 1, iters_cap: Optional[int] = None,
                  hook=None, timers: Optional[Dict[str, float]] = None) -> PairResult:
    """Registration.optimize_deformation_pyramid, registration.py:126-262, on CPU.

    RNG order matches the reference: weights first (:133-140), then two randperm (:156-157).
    `init`, `src_perm`, `tgt_perm` inject them instead (teacher forcing / reproducible pairs).
    `hook(level, it, state)` is called once per iteration before the early-stop test with the
    tensors of that iteration (used by the golden generator).  `timers`: a dict that receives the wall-clock
    seconds of the reference's timer keys lvl_warp / Chamfer / backprop (registration.py:207-213, 234-238)."""
    import time as _time
    specs = make_specs(cfg.depth, cfg.width, cfg.k0, cfg.m, cfg.rotation_format,
                       nonrigidity_est=cfg.w_reg > 0, motion=cfg.motion_type)
    if init is None:
        init = [init_params(s) for s in specs]
    params = [{k: v.detach().clone() for k, v in P.items()} for P in init]

    src_mean = src_pcd.mean(dim=0, keepdim=True)                                   # :150-153
    tgt_mean = tgt_pcd.mean(dim=0, keepdim=True)
    src_c = src_pcd - src_mean
    tgt_c = tgt_pcd - tgt_mean
    if src_perm is None:
        src_perm = torch.randperm(src_c.shape[0])                                  # :156
    if tgt_perm is None:
        tgt_perm = torch.randperm(tgt_c.shape[0])                                  # :157
    s_sample = src_c[src_perm[:cfg.samples]]
    t_sample = tgt_c[tgt_perm[:cfg.samples]]
    if landmarks is not None:
        src_ldmk = landmarks[0] - src_mean                                         # :162-164
        tgt_ldmk = landmarks[1] - tgt_mean

    iters_per_level, loss_per_level, curves = [], [], []
    n_iters = cfg.iters if iters_cap is None else min(cfg.iters, iters_cap)
    bce = torch.nn.BCELoss()
    for level, spec in enumerate(specs):
        P = params[level]
        for v in P.values():
            v.requires_grad_(True)                                                 # :173
        opt = torch.optim.Adam(list(P[n] for n, _ in param_layout(spec)), lr=cfg.lr)  # :176
        stop = EarlyStop(cfg.max_break_count, cfg.break_threshold_ratio)
        done, curve, last = 0, [], float("nan")
        for it in range(n_iters):
            if landmarks is not None:                                              # :187-203
                if cfg.w_cd > 0:
                    pts = torch.cat([src_ldmk, s_sample])
                    wp, nr = layer_forward(spec, P, pts)
                    w_ldmk, s_warped = wp[:len(src_ldmk)], wp[len(src_ldmk):]
                    loss = torch.mean(torch.sum((w_ldmk - tgt_ldmk) ** 2, dim=-1)) + cfg.w_cd * \
                        chamfer_truncated(s_warped[None], t_sample[None], trunc=cfg.trunc_cd,
                                          mode=knn_mode, threads=knn_threads)
                else:
                    w_ldmk, nr = layer_forward(spec, P, src_ldmk)
                    loss = torch.mean(torch.sum((w_ldmk - tgt_ldmk) ** 2, dim=-1))
            else:                                                                  # :205-213
                t0 = _time.perf_counter()
                s_warped, nr = layer_forward(spec, P, s_sample)
                t1 = _time.perf_counter()
                loss = chamfer_truncated(s_warped[None], t_sample[None], trunc=1e9,
                                         mode=knn_mode, threads=knn_threads)
                if timers is not None:
                    timers["lvl_warp"] = timers.get("lvl_warp", 0.0) + (t1 - t0)
                    timers["Chamfer"] = timers.get("Chamfer", 0.0) + (_time.perf_counter() - t1)
            if level > 0 and cfg.w_reg > 0:                                        # :216-220
                loss = loss + cfg.w_reg * bce(nr, torch.zeros_like(nr))
            last = loss.item()
            curve.append(last)
            if hook is not None:
                hook(level, it, dict(params=P, opt=opt, loss=loss, x_in=s_sample,
                                     x_out=s_warped if landmarks is None or cfg.w_cd > 0 else None))
            if stop.should_stop(last):                                             # :225-232
                break
            t2 = _time.perf_counter()
            opt.zero_grad()                                                        # :235-237
            loss.backward()
            opt.step()
            if timers is not None:
                timers["backprop"] = timers.get("backprop", 0.0) + (_time.perf_counter() - t2)
            done += 1
        for v in P.values():
            v.requires_grad_(False)
        if landmarks is not None:                                                  # :241-249
            src_ldmk = w_ldmk.detach()
            if cfg.w_cd > 0:
                s_sample = s_warped.detach()
        else:
            s_sample = s_warped.detach()
        iters_per_level.append(done)
        loss_per_level.append(last)
        curves.append(curve)

    with torch.no_grad():                                                          # :253-259
        warped, _ = pyramid_warp(specs, params, src_c)
        warped = warped + tgt_mean
    res = PairResult(warped, iters_per_level, loss_per_level, curves)
    res.params = params
    return res


# --------------------------------------------------------------------------------------
# Scene-flow metrics  (model/loss.py:382-403, 431-471)  -- SURVEY.md 8(f1)
# --------------------------------------------------------------------------------------
def scene_flow_metrics(pred: torch.Tensor, labels: torch.Tensor, strict=0.025, relax=0.05):
    err = torch.sqrt(torch.sum((pred - labels) ** 2, 1)).cpu()
    lab = torch.sqrt(torch.sum(labels * labels, 1)).cpu()
    rel = err / (lab + 1e-20)
    epe = torch.mean(err).item()
    acc_s = torch.mean(((err < strict) | (rel < strict)).float()).item()
    acc_r = torch.mean(((err < relax) | (rel < relax)).float()).item()
    outl = torch.mean((rel > 0.3).float()).item()
    return epe * 100, acc_s * 100, acc_r * 100, outl * 100


def compute_flow_metrics(flow, flow_gt, overlap=None):
    out = {}
    for tag, sel in (("full", None), ("vis", overlap), ("occ", None if overlap is None else ~overlap)):
        if tag != "full" and overlap is None:
            continue
        f, g = (flow, flow_gt) if sel is None else (flow[sel], flow_gt[sel])
        e, s, r, o = scene_flow_metrics(f, g)
        out.update({f"{tag}-epe": e, f"{tag}-AccS": s, f"{tag}-AccR": r, f"{tag}-outlier": o})
    return out
