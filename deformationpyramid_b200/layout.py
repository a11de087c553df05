"""Flat parameter-block layout of one pyramid level (host-side mirror of NdpLayout in
csrc/ndp_common.cuh): tensors in nn.Module.parameters() order of the reference NDPLayer
(model/nets.py:75-103)."""
from __future__ import annotations

from typing import List, Tuple

ROT_DIM = {"axis_angle": 3, "euler": 3, "quaternion": 4, "6D": 6}


def param_layout(depth: int, width: int, motion: str, rotation_format: str, nonrigidity: bool
                 ) -> List[Tuple[str, Tuple[int, ...]]]:
    W = width
    out = [("input.0.weight", (W, 6)), ("input.0.bias", (W,))]
    for l in range(depth - 1):
        out += [(f"mlp.pts_linears.{l}.weight", (W, W)), (f"mlp.pts_linears.{l}.bias", (W,))]
    if motion in ("SE3", "Sim3"):
        R = ROT_DIM[rotation_format]
        out += [("rot_brach.weight", (R, W)), ("rot_brach.bias", (R,))]
        if motion == "Sim3":
            out += [("s_branch.weight", (1, W)), ("s_branch.bias", (1,))]
    out += [("trn_branch.weight", (3, W)), ("trn_branch.bias", (3,))]
    if nonrigidity:
        out += [("nr_branch.weight", (1, W)), ("nr_branch.bias", (1,))]
    return out


def numel(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def param_count(depth: int, width: int, motion: str, rotation_format: str, nonrigidity: bool) -> int:
    return sum(numel(s) for _, s in param_layout(depth, width, motion, rotation_format, nonrigidity))
