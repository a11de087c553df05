"""Parity tests proper: the nvcc-built libndp_b200.so on a real B200, called through the C ABI
(ctypes) and through the reference-shaped Python API, checked against the golden vectors of the
unmodified reference and against the oracle.  Run with `pytest -m gpu` on the GPU box."""
import numpy as np
import pytest
import torch

from deformationpyramid_b200 import ops
from deformationpyramid_b200.synthetic import make_pair
from oracle import ndp_oracle as O
from parity_cases import (REL_TOL, rel, check_layers_against_golden, check_layers_vs_oracle_depths,
                          check_chamfer_against_golden, check_adam,
                          check_trajectory_teacher_forced, check_solver_against_oracle,
                          check_chamfer_vs_oracle_random, check_culled_search_equals_brute_force,
                          check_solver_repeatable, check_fp32_pipe_mode)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lib():
    from deformationpyramid_b200 import _lib
    return _lib.load()          # raises if the library or the GPU is missing: no fallback


def test_library_is_the_cuda_build(lib):
    assert lib.ndp_version() >= 100
    assert getattr(lib, "_ndp_requires_cuda", False)
    with pytest.raises(ValueError):
        ops.chamfer(torch.zeros(4, 3), torch.zeros(4, 3), 1e9, lib=lib)      # CPU tensors are refused


def test_layers_golden(lib, golden_dir):
    check_layers_against_golden(lib, golden_dir, DEV)


def test_backward_accumulates_over_tiles(lib):
    """Gradients of several tiles accumulated in TMEM by one CTA (what 8192-point clouds use) on small inputs."""
    from deformationpyramid_b200 import ops
    try:
        ops.set_layer_tuning(4, 0, lib=lib)
        check_layers_vs_oracle_depths(lib, DEV, cases=((3, 700), (2, 257)))
        ops.set_layer_tuning(2, 0, lib=lib)
        check_layers_vs_oracle_depths(lib, DEV, cases=((3, 385),))
        # the throughput profile of large batches: 8 tiles per backward CTA, 2 tile-pair rounds per forward CTA
        ops.set_layer_tuning(8, 2, lib=lib)
        check_layers_vs_oracle_depths(lib, DEV, cases=((3, 1300),))
    finally:
        ops.set_layer_tuning(0, 0, lib=lib)


def test_layers_other_depths_vs_oracle(lib):
    check_layers_vs_oracle_depths(lib, DEV)


def test_chamfer_golden(lib, golden_dir):
    check_chamfer_against_golden(lib, golden_dir, DEV)


def test_chamfer_random_vs_oracle(lib):
    check_chamfer_vs_oracle_random(lib, DEV, sizes=[(1, 1), (5, 700), (513, 129), (1100, 1030), (2048, 2048),
                                                    (4097, 3000)])


def test_chamfer_full_size_properties(lib):
    """BASELINE.json full size (8192 x 8192): size-independent properties + a sampled oracle check."""
    src, tgt = make_pair(0, 8192, 8192)
    x, y = src.to(DEV), tgt.to(DEV)
    loss, gx, (d2x, ix, d2y, iy) = ops.chamfer(x, y, 1e9, want_nn=True, lib=lib)
    # the reported index attains the reported distance, recomputed with the defined expression
    d = x - y[ix]
    d2 = torch.addcmul(torch.addcmul(d[:, 0] * d[:, 0], d[:, 1], d[:, 1]), d[:, 2], d[:, 2])
    assert torch.allclose(d2, d2x, rtol=1e-6, atol=0)
    # symmetric call: the roles swap exactly
    loss2, _, (e2x, jx, e2y, jy) = ops.chamfer(y, x, 1e9, want_nn=True, lib=lib)
    assert torch.equal(e2x, d2y) and torch.equal(jx, iy) and torch.equal(e2y, d2x) and torch.equal(jy, ix)
    assert abs(float(loss) - float(loss2)) <= 1e-6 * float(loss)
    # loss equals the mean NN distances
    ref = d2x.double().sqrt().mean() + d2y.double().sqrt().mean()
    assert abs(float(loss) - float(ref)) <= 1e-6 * float(ref)
    # appending far-away targets changes nothing
    far = torch.cat([y, y + 100.0])
    _, _, (f2x, kx, _, _) = ops.chamfer(x, far, 1e9, want_nn=True, lib=lib)
    assert torch.equal(kx, ix) and torch.equal(f2x, d2x)
    # 512 sampled queries against the oracle, bit-exact
    sel = torch.randperm(8192, generator=torch.Generator().manual_seed(0))[:512]
    od, oi = O.knn1(src[sel], tgt, threads=O.max_threads())
    assert torch.equal(oi, ix.cpu()[sel]) and torch.equal(od, d2x.cpu()[sel])
    # directional derivative along a rigid translation of x (a smooth direction: with an independent
    # random direction per point the NN assignments flip inside the finite-difference interval)
    v = torch.tensor([1.0, 0.5, -0.3], device=DEV).expand(8192, 3).contiguous()
    eps = 1e-3
    lp, _ = ops.chamfer((x + eps * v).contiguous(), y, 1e9, lib=lib)
    lm, _ = ops.chamfer((x - eps * v).contiguous(), y, 1e9, lib=lib)
    fd = (float(lp) - float(lm)) / (2 * eps)
    an = float((gx * v).sum())
    assert abs(fd - an) <= 2e-2 * max(abs(an), 1e-2), (fd, an)


def test_adam(lib):
    check_adam(lib, DEV)


def test_trajectory_teacher_forced(lib, golden_dir):
    check_trajectory_teacher_forced(lib, golden_dir, DEV)


def test_config2_as_written_every_step_teacher_forced(lib, golden_dir):
    from parity_cases import check_config2_teacher_forced
    check_config2_teacher_forced(lib, DEV, golden_dir)


def test_solver_config2_shape(lib):
    """BASELINE.json configs[1] shape: 2048-pt pair, single level; free-running horizon per
    SURVEY.md section 7 (hard part 3): the first iterations agree with the oracle to 1e-4."""
    check_solver_against_oracle(lib, DEV, host=False, npairs=1, n=2048, m=2048, samples=2048, levels=1, iters=8,
                                early_stop=False)


def test_solver_multi_level_batched_device(lib):
    check_solver_against_oracle(lib, DEV, host=False, npairs=3, n=700, m=650, samples=512, levels=3, iters=6,
                                early_stop=False)


def test_solver_host_buffers_early_stop_ragged(lib):
    check_solver_against_oracle(lib, DEV, host=True, npairs=2, n=150, m=140, samples=160, levels=2, iters=12,
                                early_stop=True, ratio=0.05, max_break=2)


def test_solver_sim3_euler(lib):
    check_solver_against_oracle(lib, DEV, host=False, npairs=1, n=600, m=600, samples=600, levels=2, iters=6,
                                early_stop=False, motion="Sim3", rot="euler")


def test_solver_batch_invariance(lib):
    """A pair's result does not depend on what it is batched with (deterministic kernels)."""
    cfg = dict(max_src_points=1024, max_tgt_points=1024, samples=512, levels=2, k0=-8, depth=3, width=128,
               motion="SE3", rotation_format="axis_angle", iters=20, max_break_count=3,
               break_threshold_ratio=0.01, lr=0.01, lib=lib)
    specs = O.make_specs(3, 128, -8, 2, "axis_angle")
    def mk(p):
        src, tgt = make_pair(p, 900, 800)
        torch.manual_seed(p)
        flat = torch.cat([O.flatten_params(s, O.init_params(s)) for s in specs])
        return src.to(DEV), tgt.to(DEV), flat.to(DEV)
    a, b = mk(1), mk(2)
    s1 = ops.Solver(max_pairs=1, **cfg)
    w_single, it1, _ = s1.register([a[0]], [a[1]], [a[2].clone()])
    w_again, it1b, _ = s1.register([a[0]], [a[1]], [a[2].clone()])
    s2 = ops.Solver(max_pairs=2, **cfg)
    w_batch, it2, _ = s2.register([b[0], a[0]], [b[1], a[1]], [b[2].clone(), a[2].clone()])
    assert torch.equal(w_single[0], w_again[0])                 # run-to-run bit reproducible
    assert torch.equal(w_single[0], w_batch[1]) and torch.equal(it1[0], it2[1])


def test_solver_brute_force_mode(lib):
    check_solver_against_oracle(lib, DEV, host=False, npairs=2, n=700, m=650, samples=512, levels=2, iters=6,
                                early_stop=False, nn_mode=1)


def test_culled_search_equals_brute_force(lib):
    check_culled_search_equals_brute_force(lib, DEV)
    check_culled_search_equals_brute_force(lib, DEV, n=4200, m=4100, samples=4096, levels=2, iters=12)
    check_culled_search_equals_brute_force(lib, DEV, n=8300, m=8250, samples=8192, levels=1, iters=6)   # BASELINE size


def test_full_size_regrouping_invariance(lib):
    """BASELINE.json size (8192 samples): the tensor-core path's work grouping (tiles per backward CTA,
    tile-pair rounds per forward CTA, stream groups) only regroups fp32 sums -- loss curves of a short
    run agree to rounding across groupings, and each grouping is bit-reproducible."""
    from deformationpyramid_b200 import ops
    from deformationpyramid_b200.synthetic import make_pair
    from oracle import ndp_oracle as O
    levels, iters, S = 2, 5, 8192
    specs = O.make_specs(3, 128, -8, levels, "axis_angle")
    pairs = [make_pair(90 + p, S, S) for p in range(3)]
    torch.manual_seed(1)
    flats0 = [torch.cat([O.flatten_params(s, O.init_params(s)) for s in specs]) for _ in pairs]

    def run(tpc=0, rounds=0, streams=0):
        solver = ops.Solver(max_pairs=3, max_src_points=S, max_tgt_points=S, samples=S, levels=levels, k0=-8, depth=3,
                            width=128, motion="SE3", rotation_format="axis_angle", iters=iters, max_break_count=10 ** 9,
                            break_threshold_ratio=0.001, lr=0.01, record_loss=True, tiles_per_bwd_cta=tpc,
                            fwd_rounds=rounds, streams=streams, lib=lib)
        warped, _, _ = solver.register([a.to(DEV) for a, _ in pairs], [b.to(DEV) for _, b in pairs],
                                       [f.clone().to(DEV) for f in flats0])
        out = torch.stack([solver.losses(p) for p in range(3)]), [w.cpu() for w in warped]
        solver.close()
        return out
    base, wbase = run()
    again, wagain = run()
    assert torch.equal(base, again) and all(torch.equal(a, b) for a, b in zip(wbase, wagain))
    for tpc, rounds, streams in ((1, 1, 1), (8, 2, 3), (2, 4, 2)):
        other, wother = run(tpc, rounds, streams)
        assert torch.allclose(base, other, rtol=2e-5, atol=0), (tpc, rounds, streams)
        for a, b in zip(wbase, wother):      # 10 free-running iterations amplify the regrouped fp32 sums (measured up to 1.2e-4)
            assert rel(a.numpy(), b.numpy()) < 5 * REL_TOL


def test_solver_nn_indices_bit_exact_full_size(lib):
    """Row g1 at BASELINE.json's size: all 2 x 8192 indices and squared distances of the CULLED search the solver runs,
    temporal seeds live (6 iterations), duplicated targets; then a lattice cloud (exact ties everywhere)."""
    from parity_cases import check_solver_last_nn
    check_solver_last_nn(lib, DEV, n=8300, m=8250, samples=8192, levels=2, iters=6)
    check_solver_last_nn(lib, DEV, n=8200, m=8200, samples=8192, levels=1, iters=5, lattice=True, dup=0)
    check_solver_last_nn(lib, DEV, n=700, m=650, samples=600, levels=1, iters=3, nn_mode=1)     # brute-force mode, ragged


def test_headline_config_vs_oracle(lib):
    """bench.py's default step -- 64 pairs x 8192 samples x 9 levels under the execution profile it selects (16 tiles per
    backward CTA, 8 forward rounds, 4 stream groups): pairs 0 and 63 (first / last stream group) vs the oracle; and the
    32-pair profile (8 tiles per backward CTA)."""
    from parity_cases import check_headline_config_vs_oracle
    check_headline_config_vs_oracle(lib, DEV, npairs=64, check=(0, 63))
    check_headline_config_vs_oracle(lib, DEV, npairs=32, check=(31,))


def test_solver_repeatable_with_early_stop(lib):
    check_solver_repeatable(lib, DEV)


def test_fp32_pipe_mode(lib, golden_dir):
    check_fp32_pipe_mode(lib, DEV, golden_dir)
