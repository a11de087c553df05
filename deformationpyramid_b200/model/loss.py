"""Host-side mirror of the reference model/loss.py for the NDP path.

compute_truncated_chamfer_distance keeps the reference signature (model/loss.py:94-105) and the
argument validation of :8-58, 158-164, but the two pytorch3d kNN searches, the truncation, the L1
reduction and the gradient are ONE call into kernel (2) (csrc/ndp_chamfer.cu).  CUDA only.
The flow metrics (model/loss.py:382-403, 431-471) are host-side bookkeeping of the evaluation
scripts and stay plain torch, as in the reference.
"""
from __future__ import annotations

from typing import Union

import torch

from .. import ops


def _validate_chamfer_reduction_inputs(batch_reduction: Union[str, None], point_reduction: str):
    if batch_reduction is not None and batch_reduction not in ["mean", "sum"]:
        raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
    if point_reduction not in ["mean", "sum"]:
        raise ValueError('point_reduction must be one of ["mean", "sum"]')


class _ChamferFn(torch.autograd.Function):
    """x [n,3], y [m,3] -> 0-dim loss.  Forward computes loss AND dloss/dx in the same kernel pass;
    dloss/dy (never needed on the NDP path, the target carries no gradient) is obtained from the
    symmetric call with the roles of the clouds exchanged."""

    @staticmethod
    def forward(ctx, x, y, trunc):
        xc, yc = x.detach().contiguous(), y.detach().contiguous()
        loss, gx = ops.chamfer(xc, yc, trunc)
        ctx.trunc = trunc
        ctx.save_for_backward(gx, xc, yc)
        ctx.need_y = y.requires_grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        gx, xc, yc = ctx.saved_tensors
        gy = None
        if ctx.need_y:
            _, gy = ops.chamfer(yc, xc, ctx.trunc)
            gy = gy * g
        return gx * g, gy, None


def compute_truncated_chamfer_distance(
        x,
        y,
        x_lengths=None,
        y_lengths=None,
        x_normals=None,
        y_normals=None,
        weights=None,
        trunc=0.2,
        batch_reduction: Union[str, None] = "mean",
        point_reduction: str = "mean",
):
    """Truncated L1 Chamfer distance between x (N, P1, 3) and y (N, P2, 3) -> 0-dim tensor
    (model/loss.py:94-258).  Supported: the configuration every reference call site uses
    (registration.py:195,212, shape_transfer.py:132): full lengths, no normals, no weights,
    "mean"/"mean" reductions.  Anything else raises instead of silently taking another path."""
    _validate_chamfer_reduction_inputs(batch_reduction, point_reduction)
    if not torch.is_tensor(x) or not torch.is_tensor(y):
        raise ValueError("The input pointclouds should be torch.Tensor of shape (minibatch, num_points, 3).")
    if x.ndim != 3 or y.ndim != 3:
        raise ValueError("Expected points to be of shape (N, P, D)")
    N, P1, D = x.shape
    P2 = y.shape[1]
    if y.shape[0] != N or y.shape[2] != D:
        raise ValueError("y does not have the correct shape.")
    if D != 3:
        raise ValueError("the sm_100a Chamfer kernel handles 3-D points")
    for lens, P in ((x_lengths, P1), (y_lengths, P2)):
        if lens is not None:
            if lens.ndim != 1 or lens.shape[0] != N:
                raise ValueError("Expected lengths to be of shape (N,)")
            if bool((lens != P).any()):
                raise NotImplementedError("heterogeneous lengths are not used on the NDP path")
    if x_normals is not None or y_normals is not None:
        raise NotImplementedError("normals are not used on the NDP path")
    if weights is not None:
        raise NotImplementedError("batch weights are not used on the NDP path")
    if batch_reduction != "mean" or point_reduction != "mean":
        raise NotImplementedError("only the reference's 'mean'/'mean' reductions are implemented")
    if not x.is_cuda and ops.requires_cuda():
        raise RuntimeError("compute_truncated_chamfer_distance runs on CUDA only (no CPU fallback)")
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise ValueError("float32 point clouds expected")

    total = None
    for b in range(N):
        lb = _ChamferFn.apply(x[b], y[b], float(trunc))
        total = lb if total is None else total + lb
    return total / N if N != 1 else total


# ---- evaluation metrics (host-side; model/loss.py:382-403, 431-471) ---------------------------------
def scene_flow_metrics(pred, labels, strict=0.025, relax=0.05):
    l2_norm = torch.sqrt(torch.sum((pred - labels) ** 2, 1)).cpu()
    labels_norm = torch.sqrt(torch.sum(labels * labels, 1)).cpu()
    relative_err = l2_norm / (labels_norm + 1e-20)
    EPE3D = torch.mean(l2_norm).item()
    AccS = torch.mean(((l2_norm < strict) | (relative_err < strict)).float()).item()
    AccR = torch.mean(((l2_norm < relax) | (relative_err < relax)).float()).item()
    outlier = torch.mean((relative_err > 0.3).float()).item()
    return EPE3D * 100, AccS * 100, AccR * 100, outlier * 100


def compute_flow_metrics(flow, flow_gt, overlap=None):
    metric_info = {}
    subsets = [("full", flow, flow_gt)]
    if overlap is not None:
        subsets += [("vis", flow[overlap], flow_gt[overlap]), ("occ", flow[~overlap], flow_gt[~overlap])]
    for tag, f, g in subsets:
        epe, AccS, AccR, outlier = scene_flow_metrics(f, g)
        metric_info.update({f"{tag}-epe": epe, f"{tag}-AccS": AccS, f"{tag}-AccR": AccR,
                            f"{tag}-outlier": outlier})
    return metric_info
