"""The oracle (oracle/ndp_oracle.py + oracle/knn_oracle.c) against the golden vectors produced by
the UNMODIFIED reference (oracle/gen_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ndp_oracle as O
from deformationpyramid_b200.synthetic import make_pair

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


# ---- pins added in round 2: metrics, landmark / regulariser branches, config 2 as written, mesh fixtures -------------
def test_flow_metrics_against_reference(golden_dir):
    """oracle.compute_flow_metrics / scene_flow_metrics AND the mirror's (deformationpyramid_b200.model.loss) against the
    unmodified model/loss.py:382-403, 431-471."""
    from deformationpyramid_b200.model import loss as mirror
    G = _load(golden_dir, "metrics.npz")
    for name in G["meta"]:
        flow, gt = torch.from_numpy(G[f"{name}_flow"]), torch.from_numpy(G[f"{name}_gt"])
        ov = torch.from_numpy(G[f"{name}_overlap"])
        for impl in (O, mirror):
            m = impl.compute_flow_metrics(flow, gt, overlap=ov)
            assert list(m.keys()) == [str(k) for k in G[f"{name}_keys"]]
            got, want = np.array([m[k] for k in m]), G[f"{name}_vals"]
            assert np.array_equal(np.isnan(got), np.isnan(want))
            assert np.allclose(got[~np.isnan(got)], want[~np.isnan(want)], rtol=1e-6, atol=1e-9), (name, impl.__name__)
            m2 = impl.compute_flow_metrics(flow, gt)
            assert list(m2.keys()) == [str(k) for k in G[f"{name}_keys_noov"]]
            assert np.allclose(np.array([m2[k] for k in m2]), G[f"{name}_vals_noov"], rtol=1e-6)
            assert np.allclose(np.array(impl.scene_flow_metrics(flow, gt)), G[f"{name}_scene"], rtol=1e-6)


def _branch_cfg(G, name):
    kv = dict(s.split("=", 1) for s in [str(x) for x in G[f"{name}_cfg"]])
    return O.NDPConfig(iters=int(kv["iters"]), lr=float(kv["lr"]), max_break_count=int(kv["max_break_count"]),
                       break_threshold_ratio=float(kv["break_threshold_ratio"]), w_reg=float(kv["w_reg"]),
                       samples=int(kv["samples"]), m=int(kv["m"]), k0=int(kv["k0"]), depth=int(kv["depth"]),
                       width=int(kv["width"]), motion_type=kv["motion_type"], rotation_format=kv["rotation_format"],
                       w_cd=float(kv.get("w_cd", 0.0)), trunc_cd=float(kv.get("trunc_cd", 0.25)))


@pytest.mark.parametrize("name", ["ldmk", "ldmk_cd", "wreg"])
def test_landmark_and_regulariser_branches(golden_dir, name):
    """oracle.optimize_pair's landmark (registration.py:187-203, w_cd = 0 and w_cd > 0 with truncation) and nonrigidity
    (:216-220) branches vs the unmodified Registration.register(): same Adam steps per level, Chamfer values of the
    first iterations to rounding, final cloud within the free-running horizon."""
    G = _load(golden_dir, "branches.npz")
    p, n, m, seed, n_ldmk = [int(v) for v in G[f"{name}_pair"]]
    src, tgt = make_pair(p, n, m)
    landmarks = (torch.from_numpy(G[f"{name}_ldmk_s"]), torch.from_numpy(G[f"{name}_ldmk_t"])) if n_ldmk else None
    cfg = _branch_cfg(G, name)
    torch.manual_seed(seed)
    ref = O.optimize_pair(cfg, src, tgt, landmarks=landmarks)
    assert ref.iters_per_level == [int(v) for v in G[f"{name}_steps"]]
    assert _rel(ref.warped.numpy(), G[f"{name}_warped"]) < 2e-3


def test_config2_as_written_loss_curve(golden_dir):
    """BASELINE.json config 2 as written (2048 points, one level, 200 Adam iterations, early stop off): the oracle's loss
    curve against the unmodified reference's over ALL 200 iterations."""
    G = _load(golden_dir, "config2.npz")
    p, n, m, seed = [int(v) for v in G["pair"]]
    src, tgt = make_pair(p, n, m)
    torch.manual_seed(seed)
    ref = O.optimize_pair(O.NDPConfig(m=1, iters=200, samples=2048, max_break_count=10 ** 9), src, tgt,
                          knn_threads=O.max_threads())
    mine, want = np.array(ref.loss_curve[0]), G["losses"]
    assert len(mine) == 200
    dev = np.abs(mine - want) / want
    assert dev[:20].max() < 1e-5 and dev.max() < 5e-2, (dev[:20].max(), dev.max())     # 1e-6 / 1e-4 on the generating host, 1e-7 / 1.9e-2 on the GPU box's host (chaotic after ~20 iterations)
    assert _rel(ref.warped.numpy(), G["warped"]) < 5e-2


def test_mesh_fixtures_are_the_reference_files(golden_dir):
    """tests/golden/*.ply.gz are sim3_demo/*.ply byte for byte (BASELINE.json config 1) and parse to the documented sizes."""
    import gzip
    import hashlib
    from deformationpyramid_b200.shape_transfer import read_ply_ascii
    want = {"AlienSoldier": (24856, "4e028d635b12da69"), "Ortiz": (26575, "3a02764e7bb43e0b")}
    for name, (nv, digest) in want.items():
        path = os.path.join(golden_dir, f"{name}.ply.gz")
        with gzip.open(path, "rb") as f:
            raw = f.read()
        assert hashlib.sha256(raw).hexdigest()[:16] == digest
        ref_path = f"/root/reference/sim3_demo/{name}.ply"
        if os.path.exists(ref_path):
            with open(ref_path, "rb") as f:
                assert f.read() == raw
    v, f, props = read_ply_ascii(os.path.join(golden_dir, "AlienSoldier.ply.gz"))
    assert v.shape == (24856, 3) and props[:3] == ["x", "y", "z"] and f.shape[1] == 3 and f.shape[0] >= 46108
    assert f.max() < 24856 and np.isfinite(v).all()
