"""ORACLE - TEST INFRASTRUCTURE ONLY.

Import shims that let the UNMODIFIED reference modules (model/loss.py, model/registration.py,
utils/vis.py under /root/reference) be imported in the build container, where pytorch3d,
open3d, skimage, easydict and mayavi are absent.  Only one shim carries
arithmetic: pytorch3d.ops.knn.knn_points (K=1), served by oracle/knn_oracle.c.  Everything else
is import-only.  Used by oracle/gen_golden.py; never by the product.
"""
from __future__ import annotations

import sys
import types
from collections import namedtuple

import torch

from . import ndp_oracle

_KNN = namedtuple("KNN", "dists idx knn")

KNN_MODE = 0       # 0 = fma rounding (parity contract), 1 = separately rounded (x86 pytorch3d)
KNN_THREADS = 1


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False,
               return_sorted=True):
    """pytorch3d.ops.knn.knn_points restricted to what model/loss.py:177-178 uses: K=1, full
    lengths.  Returns dists [N,P1,1] (squared), idx [N,P1,1] int64."""
    if K != 1:
        raise NotImplementedError("oracle knn_points shim supports K=1 only")
    N = p1.shape[0]
    for lens, P in ((lengths1, p1.shape[1]), (lengths2, p2.shape[1])):
        if lens is not None and bool((lens != P).any()):
            raise NotImplementedError("oracle knn_points shim supports homogeneous lengths only")
    d, i = [], []
    for b in range(N):
        db, ib = ndp_oracle.knn1_autograd(p1[b], p2[b], KNN_MODE, KNN_THREADS)
        d.append(db[:, None])
        i.append(ib[:, None])
    return _KNN(torch.stack(d), torch.stack(i), None)


def knn_gather(x, idx, lengths=None):
    N, M, U = x.shape
    K = idx.shape[2]
    return x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))


class Pointclouds:  # import-only (model/loss.py:5 uses it in an isinstance test)
    pass


class _AttrDict(dict):
    """easydict.EasyDict stand-in: attribute access on a dict (recursive for nested dicts)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def install(reference_root: str = "/root/reference") -> None:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    p3d = mod("pytorch3d")
    p3d.ops = mod("pytorch3d.ops")
    p3d.ops.knn = mod("pytorch3d.ops.knn", knn_points=knn_points, knn_gather=knn_gather)
    p3d.structures = mod("pytorch3d.structures")
    p3d.structures.pointclouds = mod("pytorch3d.structures.pointclouds", Pointclouds=Pointclouds)
    sk = mod("skimage")
    sk.io = mod("skimage.io")
    mod("open3d")
    mod("easydict", EasyDict=_AttrDict)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
