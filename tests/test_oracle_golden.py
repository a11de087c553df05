"""The oracle (oracle/ndp_oracle.py + oracle/knn_oracle.c) against the golden vectors produced by
the UNMODIFIED reference (oracle/gen_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ndp_oracle as O

TOL = 2e-6   # restatement vs reference: same fp32 math, different op grouping


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _spec(meta):
    motion, fmt, nr, m, seed, depth = meta.split(",")
    return O.LayerSpec(depth=int(depth), width=128, k0=-8, m=int(m), rotation_format=fmt,
                       nonrigidity_est=bool(int(nr)), motion=motion), int(seed)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_layer_forward_backward_all_variants(golden_dir):
    G = _load(golden_dir, "layers.npz")
    for vi, meta in enumerate(G["meta"]):
        spec, seed = _spec(str(meta))
        k = f"v{vi}"
        # constructor parity: same RNG consumption order as the reference NDPLayer
        torch.manual_seed(seed)
        fresh = O.flatten_params(spec, O.init_params(spec)).numpy()
        assert hashlib.sha256(fresh.tobytes()).hexdigest() == str(G[f"{k}_init_sha"]), meta
        assert [n for n, _ in O.param_layout(spec)] == list(G[f"{k}_names"]), meta
        assert O.param_count(spec) == G[f"{k}_params"].size

        flat = torch.from_numpy(G[f"{k}_params"]).clone().requires_grad_(True)
        P = O.unflatten_params(spec, flat)
        x = torch.from_numpy(G[f"{k}_x"]).clone().requires_grad_(True)
        y, nu = O.layer_forward(spec, P, x)
        assert _rel(y.detach().numpy(), G[f"{k}_y"]) < TOL, meta
        obj = (y * torch.from_numpy(G[f"{k}_gy"])).sum()
        if spec.nonrigidity_est:
            assert _rel(nu.detach().numpy(), G[f"{k}_nu"]) < TOL, meta
            obj = obj + (nu * torch.from_numpy(G[f"{k}_gnu"])).sum()
        else:
            assert nu is None
        gx, gp = torch.autograd.grad(obj, [x, flat])
        assert _rel(gx.numpy(), G[f"{k}_gx"]) < 2e-5, meta
        assert _rel(gp.numpy(), G[f"{k}_gparams"]) < 2e-5, meta


def test_chamfer_against_reference(golden_dir):
    G = _load(golden_dir, "chamfer.npz")
    for name in G["meta"]:
        name = str(name)
        x = torch.from_numpy(G[f"{name}_x"]).clone().requires_grad_(True)
        y = torch.from_numpy(G[f"{name}_y"])
        loss, nn = O.chamfer_truncated(x[None], y[None], trunc=float(G[f"{name}_trunc"]),
                                       return_nn=True)
        d2x, ix, d2y, iy = nn[0]
        assert np.array_equal(ix.numpy(), G[f"{name}_ix"]), name
        assert np.array_equal(iy.numpy(), G[f"{name}_iy"]), name
        assert np.array_equal(d2x.numpy(), G[f"{name}_d2x"]), name
        assert np.array_equal(d2y.numpy(), G[f"{name}_d2y"]), name
        assert abs(float(loss) - float(G[f"{name}_loss"])) <= 1e-6 * abs(float(G[f"{name}_loss"])), name
        gx, = torch.autograd.grad(loss, x)
        assert _rel(gx.numpy(), G[f"{name}_gx"]) < 1e-6, name


def test_knn_ties_take_lowest_index():
    y = torch.tensor([[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 0]])
    x = torch.tensor([[0.9, 0, 0], [0.1, 0, 0], [0.5, 0, 0]])
    d2, idx = O.knn1(x, y)
    assert idx.tolist() == [1, 0, 0]
    # threads do not change the result
    xs, ys = torch.randn(501, 3), torch.randn(333, 3)
    a = O.knn1(xs, ys, threads=1)
    b = O.knn1(xs, ys, threads=4)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_knn_against_torch_argmin():
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(400, 3, generator=g), torch.randn(300, 3, generator=g)
    d2, idx = O.knn1(x, y)
    D = ((x[:, None, :].double() - y[None].double()) ** 2).sum(-1)
    ref = D.argmin(1)
    # the fp64 argmin agrees except at fp32 near-ties; distances must agree to fp32 rounding
    assert (ref == idx).float().mean() > 0.995
    assert torch.allclose(d2.double(), D.gather(1, idx[:, None])[:, 0], rtol=1e-6, atol=0)


def test_fma_vs_separate_rounding_flip_rate():
    """Documents SURVEY.md section 3.4: the two roundings of the 3-term sum may flip near-ties."""
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2000, 3, generator=g), torch.randn(4096, 3, generator=g)
    flips = O.knn1_mode_disagreements(x, y)
    assert 0 <= flips <= 5


def test_trajectory_teacher_forced(golden_dir):
    """One oracle iteration from each recorded reference state reproduces the reference's
    warped points, loss, gradients and Adam-updated parameters."""
    G = _load(golden_dir, "trajectory.npz")
    spec = O.LayerSpec(depth=3, width=128, k0=-8, m=1)
    src, tgt = torch.from_numpy(G["src"]), torch.from_numpy(G["tgt"])
    torch.manual_seed(int(G["seed"]))
    init = O.init_params(spec)                    # registration.py:133 consumes the RNG first
    sp, tp = torch.randperm(512), torch.randperm(512)   # then :156-157
    s_sample = (src - src.mean(0, keepdim=True))[sp[:512]]
    t_sample = (tgt - tgt.mean(0, keepdim=True))[tp[:512]]
    assert np.array_equal(t_sample.numpy(), G["t_sample"])
    assert np.array_equal(O.flatten_params(spec, init).numpy(), G["it0_params_before"])
    for it in G["keep_its"]:
        k = f"it{int(it)}"
        flat = torch.from_numpy(G[f"{k}_params_before"]).clone().requires_grad_(True)
        y, _ = O.layer_forward(spec, O.unflatten_params(spec, flat), s_sample)
        assert _rel(y.detach().numpy(), G[f"{k}_x_warped"]) < TOL
        loss = O.chamfer_truncated(y[None], t_sample[None], trunc=1e9)
        assert abs(float(loss) - float(G["losses"][int(it)])) < 1e-6
        g, = torch.autograd.grad(loss, flat)
        assert _rel(g.numpy(), G[f"{k}_grads"]) < 5e-5
        # Adam from the recorded state with the REFERENCE gradient (isolates the update rule)
        p = torch.from_numpy(G[f"{k}_params_before"]).clone().requires_grad_(True)
        opt = torch.optim.Adam([p], lr=0.01)
        step = int(G[f"{k}_step_before"])
        if step > 0:
            opt.state[p] = dict(step=torch.tensor(float(step)),
                                exp_avg=torch.from_numpy(G[f"{k}_m_before"]).clone(),
                                exp_avg_sq=torch.from_numpy(G[f"{k}_v_before"]).clone())
        p.grad = torch.from_numpy(G[f"{k}_grads"]).clone()
        opt.step()
        assert _rel(p.detach().numpy(), G[f"{k}_params_after"]) < 1e-6


@pytest.mark.parametrize("name", ["ndp9", "fixed3", "sim3euler"])
def test_whole_pair_driver(golden_dir, name):
    """oracle.optimize_pair vs the unmodified Registration.register(): same iteration counts in
    the first levels, loss curves equal while the trajectories have not yet diverged
    (hard part 3), final warped cloud close."""
    from deformationpyramid_b200.synthetic import make_pair
    G = _load(golden_dir, "pairs.npz")
    p, n, m, seed = [int(v) for v in G[f"{name}_pair"]]
    kv = dict(s.split("=", 1) for s in G[f"{name}_cfg"])
    cfg = O.NDPConfig(iters=int(kv["iters"]), lr=float(kv["lr"]),
                      max_break_count=int(kv["max_break_count"]),
                      break_threshold_ratio=float(kv["break_threshold_ratio"]),
                      w_reg=float(kv["w_reg"]), samples=int(kv["samples"]), m=int(kv["m"]),
                      k0=int(kv["k0"]), depth=int(kv["depth"]), width=int(kv["width"]),
                      motion_type=kv["motion_type"], rotation_format=kv["rotation_format"])
    src, tgt = make_pair(p, n, m)
    torch.manual_seed(seed)
    res = O.optimize_pair(cfg, src, tgt)
    mine = np.array([l for c in res.loss_curve for l in c], np.float32)
    ref = G[f"{name}_losses"]
    k = min(10, len(ref), len(mine))
    assert np.allclose(mine[:k], ref[:k], rtol=1e-5, atol=1e-7)
    assert abs(len(mine) - len(ref)) <= max(3, 0.1 * len(ref))
    assert abs(mine[-1] - ref[-1]) < 0.02 * ref[-1]
    err = np.abs(res.warped.numpy() - G[f"{name}_warped"]).max()
    assert err < 2e-2
