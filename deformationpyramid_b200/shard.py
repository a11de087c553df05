"""Pair-level data parallelism and the evaluation harness around the hot path.

The reference processes pairs strictly one after another in one process (eval_nolearned.py:70).
Pairs share nothing (fresh weights, Adam state and sub-sampling per pair, registration.py:133,
156-159,176), so they shard over one process per GPU with NO data-path collective; the only
exchange is the final gather of the per-pair metric rows (NCCL on GPUs, gloo in the CPU tests).

  shard_indices        static interleave  i = r (mod R)           (the sharding point, eval_nolearned.py:70)
  FourDMatchPairs      the 4DMatch .npz wire format               (correspondence/datasets/_4dmatch.py:43-153)
  ground_truth_flow    GT scene flow + overlap mask                (eval_nolearned.py:75-84)
  gather_metric_rows   all_gather of ragged [pairs_local, C] rows  (replaces the in-process AverageMeter loop)
  evaluate             the loop of eval_nolearned.py:70-143, sharded + batched, per-pair seeded
"""
from __future__ import annotations

import glob
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

METRIC_KEYS = [f"{s}-{m}" for s in ("full", "vis", "occ") for m in ("epe", "AccS", "AccR", "outlier")]


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    if not (0 <= rank < world):
        raise ValueError("rank must be in [0, world)")
    return list(range(rank, n_items, world))


class FourDMatchPairs:
    """Reader of the 4DMatch / 4DLoMatch split directories: data_root/split/<seq>/<pair>.npz with
    keys rot, trans, s2t_flow, s_pc, t_pc, correspondences (, metric_index).  Entries are SORTED
    (the reference uses unsorted glob order, _4dmatch.py:44) so that every rank sees the same
    list.  Clouds above 30 000 points are randomly down-sampled (_4dmatch.py:30, 92-98)."""
    max_points = 30000

    def __init__(self, data_root: str, split: str):
        self.entries = sorted(glob.glob(os.path.join(data_root, split, "*/*.npz")))

    def __len__(self):
        return len(self.entries)

    def __getitem__(self, index):
        with np.load(self.entries[index]) as e:
            rot, trans, s2t_flow = e["rot"], e["trans"], e["s2t_flow"]
            src_pcd, tgt_pcd, corr = e["s_pc"], e["t_pc"], e["correspondences"]
        if src_pcd.shape[0] > self.max_points:
            src_pcd = src_pcd[np.random.permutation(src_pcd.shape[0])[:self.max_points]]
        if tgt_pcd.shape[0] > self.max_points:
            tgt_pcd = tgt_pcd[np.random.permutation(tgt_pcd.shape[0])[:self.max_points]]
        if trans.ndim == 1:
            trans = trans[:, None]
        return dict(src_pcd=src_pcd.astype(np.float32), tgt_pcd=tgt_pcd.astype(np.float32),
                    correspondences=corr, rot=rot.astype(np.float32), trans=trans.astype(np.float32),
                    s2t_flow=s2t_flow.astype(np.float32))


def ground_truth_flow(item) -> Tuple[torch.Tensor, torch.Tensor]:
    """eval_nolearned.py:75-84: flow_gt = R (src + flow) + t - src; overlap = points with a correspondence."""
    src = item["src_pcd"]
    wrapped = (item["rot"] @ (src + item["s2t_flow"]).T + item["trans"]).T
    flow_gt = torch.from_numpy((wrapped - src).astype(np.float32))
    overlap = np.zeros(len(src))
    overlap[item["correspondences"][:, 0]] = 1
    return flow_gt, torch.from_numpy(overlap.astype(bool))


def gather_metric_rows(rows: torch.Tensor, device: Optional[torch.device] = None) -> torch.Tensor:
    """rows [k_local, C] float64 (column 0 = global pair index) -> all ranks' rows, sorted by pair
    index, on every rank.  One all_gather of the counts, one of the padded rows."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows[torch.argsort(rows[:, 0])] if rows.numel() else rows
    world = dist.get_world_size()
    dev = device if device is not None else rows.device
    rows = rows.to(dev)
    cnt = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    kmax = int(max(int(c) for c in cnts))
    pad = torch.zeros(kmax, rows.shape[1], dtype=rows.dtype, device=dev)
    pad[:rows.shape[0]] = rows
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.cat([b[:int(c)] for b, c in zip(bufs, cnts)]).cpu()
    return out[torch.argsort(out[:, 0])]


def average_metrics(rows: torch.Tensor, keys: Sequence[str] = METRIC_KEYS) -> Dict[str, float]:
    """The AverageMeter reduction of eval_nolearned.py:138-152 over gathered rows (NaN rows, e.g. a
    pair without occluded points, poison the mean exactly as in the reference)."""
    return {k: float(rows[:, 1 + i].mean()) for i, k in enumerate(keys)} if rows.numel() else {}


def cloud_checksum(t: torch.Tensor) -> float:
    """48 bits of the SHA-256 of a cloud's bytes as an exactly representable float64: lets ranks compare warped
    clouds bit for bit through the metric gather."""
    import hashlib
    return float(int.from_bytes(hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).digest()[:6], "big"))


def evaluate(registration, n_items: int, get_item: Callable[[int], dict], rank: int = 0, world: int = 1,
             batch: int = 8, base_seed: int = 0, compute_metrics: bool = True, gather_device=None,
             host: bool = False, checksum: bool = False):
    """Sharded, batched version of the loop at eval_nolearned.py:70-143.

    get_item(i) -> dict(src_pcd, tgt_pcd[, correspondences, rot, trans, s2t_flow]) (numpy).
    Pair i is registered with torch.manual_seed(base_seed + i) applied before its weights and
    permutations are drawn, so the result does not depend on world size or (for a given execution
    profile, see ops.execution_profile) on the batch it runs in.  Loading and host-side preparation of
    batch k + 1 overlap the GPU work of batch k (Registration.register_batches).  host=True: the clouds
    stay (pinned) host tensors and the host<->device copies happen inside the native call.
    Returns (rows gathered on every rank [n_items, 1 + 12 (+ 1 checksum of the warped cloud)], averages)."""
    from .model.loss import compute_flow_metrics
    mine = shard_indices(n_items, rank, world)
    groups = [mine[b0:b0 + batch] for b0 in range(0, len(mine), batch)]
    loaded = []

    def batches():
        for idxs in groups:
            items = [get_item(i) for i in idxs]
            pairs = []
            for it in items:
                s = torch.from_numpy(np.ascontiguousarray(it["src_pcd"], dtype=np.float32))
                t = torch.from_numpy(np.ascontiguousarray(it["tgt_pcd"], dtype=np.float32))
                if host and torch.cuda.is_available():
                    s, t = s.pin_memory(), t.pin_memory()
                pairs.append((s, t))
            loaded.append((idxs, items, pairs))
            yield pairs

    rows = []
    seeds = ([base_seed + i for i in idxs] for idxs in groups)
    if hasattr(registration, "register_batches"):
        results = registration.register_batches(batches(), seeds=seeds, host=host)
    else:                                   # an object with only the one-batch entry point
        results = (registration.register_batch(b, seeds=sd, host=host) for b, sd in zip(batches(), seeds))
    for warped, iters, losses in results:
        idxs, items, pairs = loaded.pop(0)
        for i, it, w, (src, _), ls in zip(idxs, items, warped, pairs, losses):
            row = [float(i)]
            if compute_metrics and "s2t_flow" in it:
                flow_gt, overlap = ground_truth_flow(it)
                flow = w.detach().cpu() - src
                m = compute_flow_metrics(flow, flow_gt, overlap=overlap)
                row += [m[k] for k in METRIC_KEYS]
            else:
                row += [float(ls[-1])] + [float("nan")] * (len(METRIC_KEYS) - 1)
            if checksum:
                row.append(cloud_checksum(w))
            rows.append(row)
    width = 1 + len(METRIC_KEYS) + (1 if checksum else 0)
    rows_t = torch.tensor(rows, dtype=torch.float64).reshape(-1, width)
    allrows = gather_metric_rows(rows_t, device=gather_device)
    return allrows, average_metrics(allrows)
