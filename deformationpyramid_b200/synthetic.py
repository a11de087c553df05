"""Synthetic point-cloud pairs for benchmarks and parity tests.

Pair p is fully determined by p: a metre-scale closed surface sampled N times for the source
and, INDEPENDENTLY, M times for the target, the target pushed through a smooth non-rigid warp
(10 degree rotation about a seed-dependent axis, a small translation and a sinusoidal
displacement).  Nearest-neighbour distances are never exactly zero, so the sqrt'(0) NaN of the
reference loss (model/loss.py:227) is not triggered, and the problem is a genuine non-rigid
registration at the scale the pyramid frequencies 2^-7..2^1 rad/unit assume.
"""
from __future__ import annotations

import math

import torch


def _surface(g: torch.Generator, n: int) -> torch.Tensor:
    u = torch.randn(n, 3, generator=g, dtype=torch.float32)
    u = u / u.norm(dim=1, keepdim=True)
    r = 0.5 * (1.0 + 0.1 * torch.sin(4.0 * u[:, 0:1]))
    return (u * r).contiguous()


def make_pair(p: int, n: int = 8192, m: int = 8192):
    """Returns (src [n,3], tgt [m,3]) fp32 CPU tensors for pair index p."""
    g = torch.Generator().manual_seed(1000 + int(p))
    src = _surface(g, n)
    tgt = _surface(g, m)
    axis = torch.randn(3, generator=g, dtype=torch.float32)
    axis = axis / axis.norm()
    trn = 0.05 * torch.randn(3, generator=g, dtype=torch.float32)
    ang = math.radians(10.0)
    K = torch.tensor([[0.0, -axis[2], axis[1]], [axis[2], 0.0, -axis[0]], [-axis[1], axis[0], 0.0]])
    R = torch.eye(3) + math.sin(ang) * K + (1.0 - math.cos(ang)) * (K @ K)
    tgt = tgt @ R.T + trn
    tgt = tgt + 0.05 * torch.sin(3.0 * tgt[:, [1, 2, 0]])
    return src.contiguous(), tgt.contiguous()
