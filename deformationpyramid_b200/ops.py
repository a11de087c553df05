"""Tensor-level wrappers of the C ABI: torch owns device memory and streams, the library computes.

Every function takes an optional `lib` (default: the product library from _lib.load()); tests pass
the CPU-emulated build of the same sources together with CPU tensors.  With the product library
all tensors must live on one CUDA device; work is queued on torch's current stream.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LayerCfg, SolverCfg, MOTION, ROT_FORMAT


def _get(lib):
    return lib if lib is not None else _lib.load()


def requires_cuda() -> bool:
    """True for the product library (always: there is no CPU fallback).  Only the test suite's CPU-emulated build of
    the kernel sources, injected by tests as _lib._LIB, reports False."""
    return bool(getattr(_lib._LIB, "_ndp_requires_cuda", True))


def _stream(lib, ref: torch.Tensor):
    if getattr(lib, "_ndp_requires_cuda", False):
        return ctypes.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream)
    return ctypes.c_void_p(0)


def _chk_tensor(lib, t: torch.Tensor, name: str, dtype=torch.float32):
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if getattr(lib, "_ndp_requires_cuda", False) and not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def make_layer_cfg(depth: int, width: int, k0: int, m: int, rotation_format: str, nonrigidity: bool,
                   motion: str, mlp_scale: float = 0.001) -> LayerCfg:
    """Constructor arguments of the reference NDPLayer (model/nets.py:67) -> C struct."""
    if motion not in MOTION:
        raise AssertionError(f"motion must be one of {list(MOTION)}")          # nets.py:17
    rot = ROT_FORMAT.get(rotation_format, 0) if motion != "sflow" else 0
    if motion != "sflow" and rotation_format not in ROT_FORMAT:
        raise ValueError(f"unknown rotation_format {rotation_format!r}")
    return LayerCfg(int(width), int(depth), MOTION[motion], rot, int(bool(nonrigidity)),
                    float(2.0 ** (m + k0)), float(mlp_scale))


def param_count(cfg: LayerCfg, lib=None) -> int:
    lib = _get(lib)
    n = lib.ndp_param_count(ctypes.byref(cfg))
    if n < 0:
        raise ValueError(lib.ndp_last_error().decode())
    return int(n)


def pack_params(cfg: LayerCfg, params: torch.Tensor, pack: Optional[torch.Tensor] = None, lib=None
                ) -> torch.Tensor:
    lib = _get(lib)
    _chk_tensor(lib, params, "params")
    n = lib.ndp_pack_count(ctypes.byref(cfg))
    if n < 0:
        raise ValueError(lib.ndp_last_error().decode())
    if pack is None:
        pack = torch.empty(int(n), dtype=torch.float32, device=params.device)
    _lib.check(lib, lib.ndp_pack_params(ctypes.byref(cfg), _ptr(params), _ptr(pack), _stream(lib, params)),
               "ndp_pack_params")
    return pack


def layer_forward(cfg: LayerCfg, params: torch.Tensor, pack: torch.Tensor, x: torch.Tensor,
                  need_saved: bool = True, lib=None
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """NDPLayer.forward (model/nets.py:111-140): x [n,3] -> (y [n,3], nu [n] | None, saved | None)."""
    lib = _get(lib)
    for t, nm in ((params, "params"), (pack, "pack"), (x, "x")):
        _chk_tensor(lib, t, nm)
    if x.ndim != 2 or x.shape[1] != 3:
        raise ValueError("x must have shape [n, 3]")
    n = x.shape[0]
    y = torch.empty_like(x)
    nu = torch.empty(n, dtype=torch.float32, device=x.device) if cfg.nonrigidity else None
    saved = None
    if need_saved:
        nf = int(lib.ndp_saved_floats(ctypes.byref(cfg), n))
        saved = torch.empty(max(4, nf), dtype=torch.float32, device=x.device)
    _lib.check(lib, lib.ndp_layer_forward(ctypes.byref(cfg), _ptr(params), _ptr(pack), _ptr(x), n, _ptr(y),
                                          _ptr(nu), _ptr(saved), _stream(lib, x)), "ndp_layer_forward")
    return y, nu, saved


def layer_backward(cfg: LayerCfg, params: torch.Tensor, pack: torch.Tensor, x: torch.Tensor, saved: torch.Tensor,
                   grad_y: torch.Tensor, grad_nu: Optional[torch.Tensor] = None, need_grad_x: bool = False,
                   lib=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Backward of NDPLayer.forward: -> (dL/dparams flat [P], dL/dx [n,3] | None)."""
    lib = _get(lib)
    for t, nm in ((params, "params"), (pack, "pack"), (x, "x"), (saved, "saved"), (grad_y, "grad_y")):
        _chk_tensor(lib, t, nm)
    if grad_nu is not None:
        _chk_tensor(lib, grad_nu, "grad_nu")
    n = x.shape[0]
    P = int(lib.ndp_param_count(ctypes.byref(cfg)))
    gparams = torch.empty(P, dtype=torch.float32, device=x.device)
    gx = torch.empty_like(x) if need_grad_x else None
    wsb = int(lib.ndp_backward_workspace_bytes(ctypes.byref(cfg), n))
    ws = torch.empty(max(4, wsb // 4), dtype=torch.float32, device=x.device)
    _lib.check(lib, lib.ndp_layer_backward(ctypes.byref(cfg), _ptr(params), _ptr(pack), _ptr(x), n, _ptr(saved),
                                           _ptr(grad_y), _ptr(grad_nu), _ptr(gparams), _ptr(gx), _ptr(ws),
                                           _stream(lib, x)), "ndp_layer_backward")
    return gparams, gx


def set_mlp_mode(mode: int, lib=None) -> None:
    """0: hidden layers on the tensor cores (default), 1: FP32 pipes.  Affects later calls / solvers."""
    lib = _get(lib)
    _lib.check(lib, lib.ndp_set_mlp_mode(int(mode)), "ndp_set_mlp_mode")


def set_layer_tuning(tiles_per_bwd_cta: int = 0, fwd_rounds: int = 0, lib=None) -> None:
    """Work grouping of the tensor-core kernels behind layer_forward / layer_backward (0 = automatic)."""
    lib = _get(lib)
    _lib.check(lib, lib.ndp_set_layer_tuning(int(tiles_per_bwd_cta), int(fwd_rounds)), "ndp_set_layer_tuning")


def chamfer(x: torch.Tensor, y: torch.Tensor, trunc: float, grad_scale: float = 1.0,
            want_nn: bool = False, lib=None):
    """compute_truncated_chamfer_distance for one pair (model/loss.py:94-258).
    x [n,3], y [m,3] -> (loss [1], dloss/dx [n,3]) and, with want_nn, (d2_x, idx_x, d2_y, idx_y)."""
    lib = _get(lib)
    _chk_tensor(lib, x, "x")
    _chk_tensor(lib, y, "y")
    if x.ndim != 2 or y.ndim != 2 or x.shape[1] != 3 or y.shape[1] != 3:
        raise ValueError("x and y must have shape [n, 3]")
    n, m = x.shape[0], y.shape[0]
    if n < 1 or m < 1:
        raise ValueError("point clouds must not be empty")
    dev = x.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    gx = torch.empty_like(x)
    d2x = idxx = d2y = idxy = None
    if want_nn:
        d2x = torch.empty(n, dtype=torch.float32, device=dev)
        idxx = torch.empty(n, dtype=torch.int64, device=dev)
        d2y = torch.empty(m, dtype=torch.float32, device=dev)
        idxy = torch.empty(m, dtype=torch.int64, device=dev)
    wsb = int(lib.ndp_chamfer_workspace_bytes(n, m))
    ws = torch.empty(wsb // 8 + 1, dtype=torch.float64, device=dev)
    _lib.check(lib, lib.ndp_chamfer(_ptr(x), n, _ptr(y), m, float(trunc), float(grad_scale), _ptr(loss), _ptr(gx),
                                    _ptr(d2x), _ptr(idxx), _ptr(d2y), _ptr(idxy), _ptr(ws), _stream(lib, x)),
               "ndp_chamfer")
    if want_nn:
        return loss, gx, (d2x, idxx, d2y, idxy)
    return loss, gx


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
              step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
              cfg: Optional[LayerCfg] = None, pack: Optional[torch.Tensor] = None, lib=None) -> None:
    """torch.optim.Adam.step() on a flat parameter block, in place (model/registration.py:237)."""
    lib = _get(lib)
    for t, nm in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk_tensor(lib, t, nm)
    cfgp = ctypes.byref(cfg) if cfg is not None else None
    _lib.check(lib, lib.ndp_adam_step(cfgp, _ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                      params.numel(), int(step), float(lr), float(beta1), float(beta2), float(eps),
                                      _ptr(pack), _stream(lib, params)), "ndp_adam_step")


def execution_profile(npairs: int, samples: int = 8192) -> dict:
    """The execution profile (ndp_solver_cfg::tiles_per_bwd_cta / fwd_rounds / streams) that
    Registration.register_batch, shard.evaluate and bench.py select from the number of pairs registered
    concurrently and the number of sampled points: below 24 pairs the library's latency-oriented defaults;
    from 24 pairs the throughput profile -- four stream groups and as many 128-sample tiles per tensor-core
    CTA (backward: 1..16, forward: half of that in tile pairs) as still leaves every stream group's launch
    about 64 CTAs (8 tiles at 32 pairs x 8192 samples, 16 at 64 pairs, 2 at 32 pairs x 2000 samples): longer
    CTAs amortise the weight loads / gradient drains, the other groups' launches fill the remaining SMs.
    Only regroups work; the gradient summation grouping follows tiles_per_bwd_cta, so a pair's result is
    bit-reproducible for a given profile."""
    if npairs < 24:
        return dict(tiles_per_bwd_cta=0, fwd_rounds=0, streams=0)
    streams = 4
    tiles = (int(samples) + 127) // 128
    want = (npairs // streams) * tiles // 64
    tpc = 1
    while tpc * 2 <= min(want, 16):
        tpc *= 2
    return dict(tiles_per_bwd_cta=tpc, fwd_rounds=max(1, tpc // 2), streams=streams)


class Solver:
    """ndp_solver: the fused per-pair driver (model/registration.py:126-262), batched over pairs."""

    def __init__(self, *, max_pairs: int, max_src_points: int, max_tgt_points: int, samples: int, levels: int,
                 k0: int, depth: int, width: int, motion: str, rotation_format: str, iters: int,
                 max_break_count: int, break_threshold_ratio: float, lr: float, trunc: float = 1e9,
                 record_loss: bool = False, profile_every: int = 0, nn_mode: int = 0, mlp_mode: Optional[str] = None,
                 tiles_per_bwd_cta: int = 0, fwd_rounds: int = 0, streams: int = 0, device=None, lib=None):
        """mlp_mode: None = process default, "tensor" = tcgen05 tensor cores, "fp32" = FP32 pipes.
        tiles_per_bwd_cta / fwd_rounds / streams: execution profile (0 = library default), see
        include/ndp_b200.h.  device: the CUDA device the solver lives on (default: torch's current device)."""
        self.lib = _get(lib)
        if mlp_mode not in (None, "tensor", "fp32"):
            raise ValueError("mlp_mode must be None, 'tensor' or 'fp32'")
        if motion not in MOTION:
            raise AssertionError(f"motion must be one of {list(MOTION)}")
        self.cfg = SolverCfg(int(max_pairs), int(max_src_points), int(max_tgt_points), int(samples), int(levels),
                             int(k0), int(depth), int(width), MOTION[motion],
                             ROT_FORMAT.get(rotation_format, 0), int(iters),
                             int(min(max_break_count, 2 ** 31 - 1)), float(break_threshold_ratio), float(lr),
                             float(trunc), int(bool(record_loss)), int(profile_every), int(nn_mode),
                             {None: 0, "tensor": 1, "fp32": 2}[mlp_mode], int(tiles_per_bwd_cta), int(fwd_rounds),
                             int(streams))
        h = ctypes.c_void_p(0)
        cuda = getattr(self.lib, "_ndp_requires_cuda", False)
        if cuda and device is not None:
            with torch.cuda.device(torch.device(device) if not isinstance(device, int) else device):
                rc = self.lib.ndp_solver_create(ctypes.byref(self.cfg), ctypes.byref(h))
        else:
            rc = self.lib.ndp_solver_create(ctypes.byref(self.cfg), ctypes.byref(h))
        _lib.check(self.lib, rc, "ndp_solver_create")
        self.handle = h
        self.params_per_pair = int(self.lib.ndp_solver_params_per_pair(h))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.ndp_solver_destroy(self.handle)
            self.handle = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.ndp_solver_launch_count(self.handle))

    @staticmethod
    def _ptr_array(ts: Optional[Sequence[Optional[torch.Tensor]]], n: int):
        if ts is None:
            return None
        arr = (ctypes.c_void_p * n)()
        for i, t in enumerate(ts):
            arr[i] = t.data_ptr() if t is not None else None
        return arr

    def register(self, src: Sequence[torch.Tensor], tgt: Sequence[torch.Tensor], params: Sequence[torch.Tensor],
                 src_perm: Optional[Sequence[torch.Tensor]] = None, tgt_perm: Optional[Sequence[torch.Tensor]] = None,
                 host: bool = False, src_samples: Optional[Sequence[int]] = None,
                 tgt_samples: Optional[Sequence[int]] = None, params_out: bool = True):
        """src[p] [ns,3], tgt[p] [nt,3], params[p] flat [levels*P] (updated in place), perms int32.
        src_samples / tgt_samples: points optimised per pair (default min(samples, n)); a permutation must hold
        exactly that many indices (its length is checked here, the C ABI receives the counts explicitly).
        host=True: `params` may also be ONE contiguous (ideally pinned) [npairs, levels*P] tensor, used in place;
        params_out=False skips the read-back of the optimised weights (the evaluation loop never looks at them).
        host=True: all tensors are (pinned) CPU tensors and the copies run inside the call.
        Returns (warped list, iters [npairs, levels] int32, last loss [npairs, levels])."""
        lib = self.lib
        npairs = len(src)
        cuda = getattr(lib, "_ndp_requires_cuda", False)
        stacked = host and torch.is_tensor(params)
        if stacked:
            if params.ndim != 2 or params.shape[0] != npairs or not params.is_contiguous():
                raise ValueError("a stacked params tensor must be contiguous [npairs, levels * param_count]")
        plist = list(params) if stacked else params
        for group, nm, dt in ((src, "src", torch.float32), (tgt, "tgt", torch.float32), (plist, "params", torch.float32),
                              (src_perm or [], "src_perm", torch.int32), (tgt_perm or [], "tgt_perm", torch.int32)):
            for t in group:
                if t.dtype != dt or not t.is_contiguous():
                    raise ValueError(f"{nm} tensors must be contiguous {dt}")
                if cuda and (t.is_cuda == host):
                    raise ValueError(f"{nm} must be {'CPU' if host else 'CUDA'} tensors for this entry point")
        for p in plist:
            if p.numel() != self.params_per_pair:
                raise ValueError("params[p] must hold levels * param_count floats")
        ns = (ctypes.c_int32 * npairs)(*[int(t.shape[0]) for t in src])
        nt = (ctypes.c_int32 * npairs)(*[int(t.shape[0]) for t in tgt])
        S = int(self.cfg.samples)
        cnt_s = [min(S, int(t.shape[0])) for t in src] if src_samples is None else [int(v) for v in src_samples]
        cnt_t = [min(S, int(t.shape[0])) for t in tgt] if tgt_samples is None else [int(v) for v in tgt_samples]
        for perms, cnts, nm in ((src_perm, cnt_s, "src_perm"), (tgt_perm, cnt_t, "tgt_perm")):
            if perms is not None:
                if len(perms) != npairs:
                    raise ValueError(f"{nm} must hold one permutation per pair")
                for t, k in zip(perms, cnts):
                    if t.numel() < k:
                        raise ValueError(f"{nm}: a permutation holds {t.numel()} indices, {k} samples are optimised")
        a_cs, a_ct = (ctypes.c_int32 * npairs)(*cnt_s), (ctypes.c_int32 * npairs)(*cnt_t)
        self._last_counts = (cnt_s, cnt_t)
        dev = src[0].device
        warped = [torch.empty_like(t) for t in src]
        iters = torch.zeros(npairs, self.cfg.levels, dtype=torch.int32)
        loss = torch.zeros(npairs, self.cfg.levels, dtype=torch.float32)
        stream = _stream(lib, src[0]) if not host else (
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream) if cuda else ctypes.c_void_p(0))
        a_src, a_tgt = self._ptr_array(src, npairs), self._ptr_array(tgt, npairs)
        a_ps, a_pt = self._ptr_array(src_perm, npairs), self._ptr_array(tgt_perm, npairs)
        a_w = self._ptr_array(warped, npairs)
        if host:
            flat = params if stacked else torch.stack([p.reshape(-1) for p in params]).contiguous()
            rc = lib.ndp_solver_register_host(self.handle, npairs, a_src, ns, a_tgt, nt, a_ps, a_pt, a_cs, a_ct, _ptr(flat),
                                              1 if params_out else 0, a_w, _ptr(iters), _ptr(loss), stream)
            _lib.check(lib, rc, "ndp_solver_register_host")
            if params_out and not stacked:
                for p, f in zip(params, flat):
                    p.copy_(f.view_as(p))
        else:
            a_par = self._ptr_array(params, npairs)
            rc = lib.ndp_solver_register_device(self.handle, npairs, a_src, ns, a_tgt, nt, a_ps, a_pt, a_cs, a_ct, a_par, a_w,
                                                _ptr(iters), _ptr(loss), stream)
            _lib.check(lib, rc, "ndp_solver_register_device")
        return warped, iters, loss

    KERNELS = ("warp_fwd", "nn_search", "chamfer_epilogue", "warp_bwd", "reduce_adam")

    def profile(self):
        """Sampled device time: ({kernel: accumulated ms}, number of sampled iterations)."""
        ms = (ctypes.c_double * 5)()
        n = ctypes.c_int64(0)
        _lib.check(self.lib, self.lib.ndp_solver_profile(self.handle, ms, ctypes.byref(n)), "ndp_solver_profile")
        return dict(zip(self.KERNELS, [float(v) for v in ms])), int(n.value)

    def nn_stats(self):
        """(distance evaluations issued by the culled search, 32-query blocks searched, most blocks one warp
        scanned in a search) since creation."""
        a, b, c = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self.lib, self.lib.ndp_solver_nn_stats(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)),
                   "ndp_solver_nn_stats")
        return int(a.value), int(b.value), int(c.value)

    @property
    def profiled_pairs(self) -> int:
        """Pairs per sampled launch (the driver runs two half-batches on two streams)."""
        return int(self.lib.ndp_solver_profiled_pairs(self.handle))

    def last_nn(self, pair: int):
        """The two K=1 searches of the last loss evaluation of the last register call (model/loss.py:177-181), in
        the sample order of that call: (idx_x int64 [n], d2_x [n], idx_y int64 [m], d2_y [m], warped source samples
        [n,3], target samples [m,3]) as CPU tensors."""
        n, m = self._last_counts[0][pair], self._last_counts[1][pair]
        idx_x, d2_x = torch.empty(n, dtype=torch.int64), torch.empty(n, dtype=torch.float32)
        idx_y, d2_y = torch.empty(m, dtype=torch.int64), torch.empty(m, dtype=torch.float32)
        w, tg = torch.empty(n, 3, dtype=torch.float32), torch.empty(m, 3, dtype=torch.float32)
        _lib.check(self.lib, self.lib.ndp_solver_last_nn(self.handle, int(pair), _ptr(idx_x), _ptr(d2_x), _ptr(idx_y),
                                                         _ptr(d2_y), _ptr(w), _ptr(tg), ctypes.c_void_p(0)),
                   "ndp_solver_last_nn")
        return idx_x, d2_x, idx_y, d2_y, w, tg

    def losses(self, pair: int) -> torch.Tensor:
        out = torch.full((self.cfg.levels, self.cfg.iters), float("nan"), dtype=torch.float32)
        _lib.check(self.lib, self.lib.ndp_solver_losses(self.handle, int(pair), _ptr(out), ctypes.c_void_p(0)),
                   "ndp_solver_losses")
        return out
