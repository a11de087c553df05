"""N>1 host logic on CPU: world_size-2 gloo runs of the sharded evaluation loop (pair sharding, per-pair
seeding, ragged metric gather) against the single-process result.  The registration object is a
deterministic stand-in -- the compute path itself is covered by the GPU tests."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from deformationpyramid_b200 import shard


class FakeRegistration:
    """register_batch with the product signature; the 'registration' is a seed-dependent rigid shift,
    so results depend on per-pair seeds exactly like fresh random weights do."""

    def register_batch(self, pairs, seeds=None, host=False):
        warped, iters, losses = [], [], []
        for (src, tgt), s in zip(pairs, seeds):
            torch.manual_seed(int(s))
            shift = 0.01 * torch.randn(3)
            warped.append(src + (tgt.mean(0) - src.mean(0)) + shift)
            iters.append(torch.tensor([int(s) % 7]))
            losses.append(torch.tensor([float(shift.norm())]))
        return warped, torch.stack(iters), torch.stack(losses)


def make_item(i):
    g = np.random.default_rng(100 + i)
    n = 50 + 3 * i
    src = g.normal(size=(n, 3)).astype(np.float32)
    flow = 0.05 * g.normal(size=(n, 3)).astype(np.float32)
    rot = np.eye(3, dtype=np.float32)
    trans = g.normal(size=(3, 1)).astype(np.float32) * 0.1
    tgt = ((rot @ (src + flow).T) + trans).T.astype(np.float32)
    corr = np.stack([np.arange(0, n, 2), np.arange(0, n, 2)], 1)
    return dict(src_pcd=src, tgt_pcd=tgt, correspondences=corr, rot=rot, trans=trans, s2t_flow=flow)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, avg = shard.evaluate(FakeRegistration(), n_items, make_item, rank=rank, world=world, batch=2, base_seed=5)
    torch.save((rows, avg), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _real_worker(rank, world, port, n_items, out_dir):
    """The REAL Registration (fused driver, register_batches pipeline) over the CPU-emulated build of the kernel sources."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from deformationpyramid_b200 import _lib
    from deformationpyramid_b200.config import ndp_config
    from deformationpyramid_b200.model.registration import Registration
    from emu_util import emu_lib
    _lib._LIB = emu_lib()
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    reg = Registration(ndp_config(samples=48, m=1, iters=2, device="cpu"))
    rows, avg = shard.evaluate(reg, n_items, make_item, rank=rank, world=world, batch=2, base_seed=5, checksum=True)
    torch.save((rows, avg), os.path.join(out_dir, f"real{world}_{rank}.pt"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_two_rank_gloo_real_registration_matches_single_process():
    """SURVEY.md section 4, T4 on the host side: identical per-pair outputs (warped-cloud checksums bit for bit) and
    identical gathered metrics whether 1 or 2 ranks register the pairs -- with the real Registration object."""
    n_items = 4
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_real_worker, args=(2, _free_port(), n_items, d), nprocs=2, join=True)
        _real_worker(0, 1, 0, n_items, d)
        r0, a0 = torch.load(os.path.join(d, "real2_0.pt"))
        r1, a1 = torch.load(os.path.join(d, "real2_1.pt"))
        rs, as_ = torch.load(os.path.join(d, "real1_0.pt"))
    assert r0.shape == (n_items, 14) and torch.equal(r0, r1)
    assert torch.equal(r0[:, -1], rs[:, -1])                        # the warped clouds, bit for bit
    assert torch.equal(torch.nan_to_num(r0), torch.nan_to_num(rs)) and a0.keys() == as_.keys()


def test_shard_indices_cover_everything_once():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in shard.shard_indices(11, r, world))
        assert seen == list(range(11))
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_two_rank_gloo_matches_single_process():
    n_items = 7                      # odd: ranks get 4 and 3 pairs (ragged gather)
    rows1, avg1 = shard.evaluate(FakeRegistration(), n_items, make_item, rank=0, world=1, batch=3, base_seed=5)
    assert rows1.shape == (n_items, 13) and rows1[:, 0].tolist() == list(range(n_items))
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_items, d), nprocs=2, join=True)
        r0, a0 = torch.load(os.path.join(d, "r0.pt"))
        r1, a1 = torch.load(os.path.join(d, "r1.pt"))
    assert torch.equal(r0, r1)                                  # every rank holds the full table
    assert torch.allclose(r0, rows1, rtol=0, atol=0, equal_nan=True)   # independent of world and batch size
    for k in avg1:
        assert (np.isnan(avg1[k]) and np.isnan(a0[k])) or abs(avg1[k] - a0[k]) < 1e-12


def test_ground_truth_flow_and_metrics_follow_reference_formulas():
    from deformationpyramid_b200.model.loss import compute_flow_metrics
    from oracle import ndp_oracle as O
    it = make_item(3)
    flow_gt, overlap = shard.ground_truth_flow(it)
    assert overlap.dtype == torch.bool and int(overlap.sum()) == len(it["correspondences"])
    pred = flow_gt + 0.01 * torch.randn_like(flow_gt)
    mine = compute_flow_metrics(pred, flow_gt, overlap=overlap)
    ref = O.compute_flow_metrics(pred, flow_gt, overlap=overlap)
    assert list(mine.keys()) == shard.METRIC_KEYS
    for k in ref:
        assert abs(mine[k] - ref[k]) < 1e-9


def test_4dmatch_reader_roundtrip(tmp_path):
    seq = tmp_path / "4DMatch-F" / "seqA"
    seq.mkdir(parents=True)
    items = [make_item(i) for i in (2, 1)]
    for name, it in zip(("cam1_0002_cam2_0004", "cam1_0000_cam2_0001"), items):
        np.savez(seq / f"{name}.npz", rot=it["rot"], trans=it["trans"][:, 0], s2t_flow=it["s2t_flow"],
                 s_pc=it["src_pcd"], t_pc=it["tgt_pcd"], correspondences=it["correspondences"])
    D = shard.FourDMatchPairs(str(tmp_path), "4DMatch-F")
    assert len(D) == 2 and D.entries == sorted(D.entries)
    got = D[0]                                       # sorted: cam1_0000... first = items[1]
    assert np.array_equal(got["src_pcd"], items[1]["src_pcd"]) and got["trans"].shape == (3, 1)
