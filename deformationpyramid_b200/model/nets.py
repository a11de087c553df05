"""Host-side mirror of the reference model/nets.py API (Deformation_Pyramid, NDPLayer, MLP).

Same constructor signatures, attribute and sub-module names (so parameters() order, state_dict keys
and the consumption of torch's RNG at construction are identical to the reference,
model/nets.py:10-62, 66-183, 295-304), but NDPLayer.forward dispatches to the sm_100a kernels
through a torch.autograd.Function: kernel (1) forward, kernel (3a) backward.  Parameters stay
ordinary nn.Parameters, so shape_transfer.py's stock torch.optim.Adam keeps working
(shape_transfer.py:121,151-153).  There is no CPU implementation: calling a layer that is not on
a CUDA device raises.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..layout import numel, param_layout


class MLP(nn.Module):
    """Container with the reference's sub-module names (nets.py:295-304); evaluated by the kernel."""

    def __init__(self, depth: int, width: int):
        super().__init__()
        self.pts_linears = nn.ModuleList([nn.Linear(width, width) for _ in range(depth - 1)])

    def forward(self, x):  # pragma: no cover - not on the product path
        raise RuntimeError("MLP is evaluated inside NDPLayer's fused kernel; call the NDPLayer")


class _NDPLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer: "NDPLayer", need_grad: bool, x: torch.Tensor, *params: torch.Tensor):
        flat = layer._flat_view(params)
        pack = ops.pack_params(layer._cfg, flat, layer._pack_buffer(flat))
        layer._pack = pack
        xc = x.detach().contiguous()
        y, nu, saved = ops.layer_forward(layer._cfg, flat, pack, xc, need_saved=need_grad)
        ctx.layer = layer
        ctx.x_needs_grad = x.requires_grad
        if need_grad:
            # the flat block may be updated in place by the optimiser before backward is called
            # a second time; autograd's version counter on `flat` is not tracked here, matching
            # the usage in shape_transfer.py (one backward per forward).
            ctx.save_for_backward(flat, xc, saved, pack)
        if nu is None:
            nu = x.new_empty(0)
        return y, nu

    @staticmethod
    def backward(ctx, gy, gnu):
        layer = ctx.layer
        flat, xc, saved, pack = ctx.saved_tensors
        gy = gy.contiguous()
        gnu_t = gnu.contiguous() if (layer.nonrigidity_est and gnu is not None) else None
        gflat, gx = ops.layer_backward(layer._cfg, flat, pack, xc, saved, gy, gnu_t, need_grad_x=ctx.x_needs_grad)
        grads, off = [], 0
        for _, shape in layer._layout:
            k = numel(shape)
            grads.append(gflat[off:off + k].view(shape))
            off += k
        return (None, None, gx) + tuple(grads)


class NDPLayer(nn.Module):
    def __init__(self, depth, width, k0, m, rotation_format="euler", nonrigidity_est=False, motion='SE3'):
        super().__init__()
        self.k0 = k0
        self.m = m
        dim_x = 6
        self.nonrigidity_est = bool(nonrigidity_est)
        self.motion = motion
        # construction order = RNG consumption order of the reference (nets.py:75-101)
        self.input = nn.Sequential(nn.Linear(dim_x, width), nn.ReLU())
        self.mlp = MLP(depth=depth, width=width)
        self.rotation_format = rotation_format
        if self.motion in ["Sim3", "SE3"]:
            rdim = {"axis_angle": 3, "euler": 3, "quaternion": 4, "6D": 6}.get(rotation_format)
            if rdim is not None:
                self.rot_brach = nn.Linear(width, rdim)        # (sic) reference spelling, nets.py:85
            if self.motion == "Sim3":
                self.s_branch = nn.Linear(width, 1)
        self.trn_branch = nn.Linear(width, 3)
        if self.nonrigidity_est:
            self.nr_branch = nn.Linear(width, 1)
            self.sigmoid = nn.Sigmoid()
        self.mlp_scale = 0.001
        self._reset_parameters()

        self.depth, self.width = depth, width
        self._cfg = ops.make_layer_cfg(depth, width, k0, m, rotation_format, self.nonrigidity_est, motion,
                                       self.mlp_scale)
        self._layout = param_layout(depth, width, motion, rotation_format, self.nonrigidity_est)
        names = [n for n, _ in self.named_parameters()]
        assert names == [n for n, _ in self._layout], (names, self._layout)
        self._flat: Optional[torch.Tensor] = None
        self._pack: Optional[torch.Tensor] = None

    def _reset_parameters(self):
        for p in self.parameters():                              # nets.py:180-183
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    # ---- flat parameter block ---------------------------------------------------------------
    def flatten_parameters_(self) -> torch.Tensor:
        """Re-home every parameter as a view of ONE contiguous fp32 block (parameters() order) so
        the kernels read the weights in place.  Values, names and Parameter objects are unchanged."""
        ps = list(self.parameters())
        dev = ps[0].device
        flat = torch.empty(sum(p.numel() for p in ps), dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            off += k
        self._flat = flat
        return flat

    def _flat_view(self, params) -> torch.Tensor:
        flat = self._flat
        if flat is not None and flat.device == params[0].device:
            off, ok, base, esz = 0, True, flat.data_ptr(), 4
            for p in params:
                if p.data_ptr() != base + off * esz or not p.is_contiguous() or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel()
            if ok:
                return flat
        # parameters were moved / replaced (e.g. by .to()): gather them (plumbing, not compute)
        return torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params]).contiguous()

    def _pack_buffer(self, flat: torch.Tensor) -> Optional[torch.Tensor]:
        if self._pack is None or self._pack.device != flat.device:
            self._pack = None
            return None
        return self._pack

    def forward(self, x):
        params = tuple(self.parameters())
        if not params[0].is_cuda and ops.requires_cuda():
            raise RuntimeError("NDPLayer runs on CUDA only (sm_100a kernels, no CPU fallback); "
                               "move the layer and its input to a CUDA device")
        if x.ndim != 2 or x.shape[-1] != 3:
            raise ValueError("NDPLayer expects points of shape [N, 3]")
        if x.dtype != torch.float32:
            raise ValueError("NDPLayer expects float32 points")
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        y, nu = _NDPLayerFn.apply(self, need_grad, x, *params)
        nonrigidity = nu if self.nonrigidity_est else None       # nets.py:132-137
        return y.squeeze(), nonrigidity

    def posenc(self, pos):
        """nets.py:164-177 (kept for API completeness; the kernel computes it inline)."""
        mul_term = 2 ** (self.m + self.k0)
        return torch.stack([torch.sin(pos * mul_term), torch.cos(pos * mul_term)], dim=-1).reshape(*pos.shape[:-1], 6)


class Deformation_Pyramid():

    def __init__(self, depth, width, device, k0, m, rotation_format, nonrigidity_est=False, motion='SE3'):
        assert motion in ["Sim3", "SE3", "sflow"]                # nets.py:17
        pyramid = []
        for i in range(m):
            layer = NDPLayer(depth, width, k0, i + 1, rotation_format,
                             nonrigidity_est=nonrigidity_est & (i != 0), motion=motion).to(device)
            layer.flatten_parameters_()
            pyramid.append(layer)
        self.pyramid = pyramid
        self.n_hierarchy = m

    def warp(self, x, max_level=None, min_level=0):
        if max_level is None:
            max_level = self.n_hierarchy - 1
        assert max_level < self.n_hierarchy, "more level than defined"
        data = {}
        for i in range(min_level, max_level + 1):
            x, nonrigidity = self.pyramid[i](x)
            data[i] = (x, nonrigidity)
        return x, data

    def gradient_setup(self, optimized_level):
        assert optimized_level < self.n_hierarchy, "more level than defined"
        # optimize current level, freeze the other levels
        for i in range(self.n_hierarchy):
            req = (i == optimized_level)
            for param in self.pyramid[i].parameters():
                param.requires_grad = req

    # ---- helpers used by Registration's fused path ----------------------------------------------
    def flat_parameters(self) -> torch.Tensor:
        """All levels back to back (level 0 first), one contiguous copy."""
        return torch.cat([layer._flat_view(tuple(layer.parameters())).reshape(-1) for layer in self.pyramid])

    def load_flat_parameters(self, flat: torch.Tensor) -> None:
        off = 0
        with torch.no_grad():
            for layer in self.pyramid:
                for p in layer.parameters():
                    k = p.numel()
                    p.copy_(flat[off:off + k].view(p.shape))
                    off += k
