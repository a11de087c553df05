"""Host-side mirror of the reference model/registration.py for deformation_model == "NDP".

Registration keeps the reference surface (model/registration.py:24-34, 93-123): attributes
src_pcd / tgt_pcd / device / config / deformation_model / landmarks, load_pcds(), register(),
optimize_deformation_pyramid(visualize=False, timer=None) -> (warped_pcd, {}, timer).

Two execution routes, both on the sm_100a kernels:
  * fused  -- landmarks is None and w_reg == 0 (config/NDP.yaml): the whole level / iteration
              loop of registration.py:170-249 runs inside the native solver (ndp_solver_*), with the
              early-stop rule evaluated on the device instead of three .item() syncs per iteration;
  * stepwise -- landmarks (LNDP, registration.py:187-203) or the nonrigidity regulariser
              (:216-220): the reference's Python loop over autograd-capable CUDA ops
              (NDPLayer / compute_truncated_chamfer_distance) and torch.optim.Adam.
The comparison methods of the paper (Nerfies, ED/N-ICP, NSFP, Sinkhorn) are out of scope.
RNG order per pair is the reference's: weights first (:133-140), then two randperm (:156-157).
"""
from __future__ import annotations

import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from .. import ops
from .loss import compute_truncated_chamfer_distance
from .nets import Deformation_Pyramid

BCE = nn.BCELoss()


def _cfg_get(config, key, default=None):
    try:
        return getattr(config, key)
    except (AttributeError, KeyError):
        return default


class Registration():

    def __init__(self, config):
        self.tgt_pcd = None
        self.src_pcd = None
        self.device = config.device
        self.config = config
        self.deformation_model = config.deformation_model
        self.landmarks = None
        self._solver = None
        self._solver_key = None
        self.NDP = None
        self.last_iters = None
        self.last_losses = None

    def load_pcds(self, src, tgt, landmarks=None):
        if type(src) in [np.ndarray]:
            src = torch.from_numpy(src)
            tgt = torch.from_numpy(tgt)
        self.src_pcd = src.to(self.device)
        self.tgt_pcd = tgt.to(self.device)
        self.landmarks = landmarks

    def register(self, **kwargs):
        if self.deformation_model == "NDP":  # deformation pyramid
            return self.optimize_deformation_pyramid(**kwargs)
        if self.deformation_model in ("Sinkhorn", "ED", "NSFP", "Nerfies"):
            raise NotImplementedError(
                f"deformation_model {self.deformation_model!r} is one of the paper's comparison methods "
                "(model/registration.py:265-572) and is outside the scope of the B200 NDP path")
        raise KeyError()

    # ------------------------------------------------------------------------------------------
    def _get_solver(self, npairs: int, ns: int, nt: int, profile_every: Optional[int] = None,
                    nn_mode: Optional[int] = None, samples: Optional[int] = None) -> ops.Solver:
        c = self.config
        samples = int(c.samples if samples is None else samples)
        nn_mode = int(_cfg_get(c, "nn_mode", 0) or 0) if nn_mode is None else int(nn_mode)
        profile = ops.execution_profile(npairs, samples)
        for k in profile:                                     # explicit config keys win over the batch-size rule
            v = _cfg_get(c, k, None)
            if v is not None:
                profile[k] = int(v)
        prof_every = int(_cfg_get(c, "profile_every", 0) or 0) if profile_every is None else int(profile_every)
        key = (samples, nn_mode, c.m, c.k0, c.depth, c.width, c.motion_type, c.rotation_format, c.iters,
               c.max_break_count, c.break_threshold_ratio, c.lr, str(self.src_pcd.device), tuple(sorted(profile.items())),
               prof_every)
        s = self._solver
        if (s is None or self._solver_key != key or s.cfg.max_pairs < npairs or s.cfg.max_src_points < ns
                or s.cfg.max_tgt_points < nt):
            if s is not None:
                s.close()
            cap = lambda v: int(max(v, 1024) * 1.25)
            self._solver = ops.Solver(max_pairs=npairs, max_src_points=cap(ns), max_tgt_points=cap(nt),
                                      samples=samples, levels=c.m, k0=c.k0, depth=c.depth, width=c.width,
                                      motion=c.motion_type, rotation_format=c.rotation_format, iters=c.iters,
                                      max_break_count=c.max_break_count,
                                      break_threshold_ratio=c.break_threshold_ratio, lr=c.lr, trunc=1e9,
                                      nn_mode=nn_mode, profile_every=prof_every,
                                      device=self.src_pcd.device, **profile)
            self._solver_key = key
        return self._solver

    def _build_pyramid(self):
        config = self.config
        return Deformation_Pyramid(depth=config.depth, width=config.width, device=self.device, k0=config.k0,
                                   m=config.m, nonrigidity_est=config.w_reg > 0,
                                   rotation_format=config.rotation_format, motion=config.motion_type)

    def optimize_deformation_pyramid(self, visualize=False, timer=None):
        config = self.config
        if visualize:
            raise NotImplementedError("visualisation (mayavi, utils/vis.py) is outside the scope of this package")
        if self.landmarks is not None and config.w_reg == 0 and float(_cfg_get(config, "w_cd", 0.0) or 0.0) == 0.0:
            return self._optimize_landmarks_fused(timer)     # config/LNDP.yaml as shipped: landmark term only
        if self.landmarks is not None or config.w_reg > 0:
            return self._optimize_stepwise(timer)

        NDP = self._build_pyramid()                              # consumes the RNG first (:133-140)
        self.src_pcd = self.src_pcd.to(self.device)
        src, tgt = self.src_pcd.contiguous(), self.tgt_pcd.contiguous()
        if src.dtype != torch.float32 or tgt.dtype != torch.float32:
            raise ValueError("float32 point clouds expected")
        sp = torch.randperm(src.shape[0])[:config.samples].to(torch.int32)     # :156-159
        tp = torch.randperm(tgt.shape[0])[:config.samples].to(torch.int32)
        dev = src.device
        flat = NDP.flat_parameters()
        # with a timer: sample the kernels of every 8th iteration with CUDA events (ndp_solver_profile) and report them
        # under the reference's keys (registration.py:207-213, 234-238)
        solver = self._get_solver(1, src.shape[0], tgt.shape[0], profile_every=8 if timer else None)
        prof0 = solver.profile() if timer else None
        if timer: timer.tic("ndp_fused")
        warped, iters, losses = solver.register([src], [tgt], [flat], [sp.to(dev)], [tp.to(dev)])
        if timer: timer.toc("ndp_fused")
        if timer: _feed_timer(timer, solver, prof0, int(iters.sum()))
        NDP.load_flat_parameters(flat)
        NDP.gradient_setup(optimized_level=-1)
        self.NDP, self.last_iters, self.last_losses = NDP, iters[0], losses[0]
        iter_cnt = {}
        return warped[0], iter_cnt, timer

    def _optimize_landmarks_fused(self, timer=None):
        """LNDP with the weights of config/LNDP.yaml (w_cd = 0, w_reg = 0): only the landmark term
        mean_i |warp(src_ldmk_i) - tgt_ldmk_i|^2 of registration.py:187-203 drives the pyramid.  Runs in the fused
        driver in paired-sample mode (ndp_solver_cfg::nn_mode = 2): the "samples" are the landmark points, source
        sample i is matched to target sample i, the early-stop rule is evaluated on the device -- no host sync per
        iteration.  The solver centres the clouds it is given on their own means; the reference centres the
        landmarks on the means of the FULL clouds (:150-153, :163-165), so centred clouds with a balancing point
        (mean exactly as computed = 0 to rounding) are passed, as in shape_transfer._register_leading_samples."""
        config = self.config
        NDP = self._build_pyramid()                              # consumes the RNG first (:133-140)
        self.src_pcd = self.src_pcd.to(self.device)
        src, tgt = self.src_pcd.contiguous(), self.tgt_pcd.contiguous()
        if src.dtype != torch.float32 or tgt.dtype != torch.float32:
            raise ValueError("float32 point clouds expected")
        torch.randperm(src.shape[0]); torch.randperm(tgt.shape[0])      # :156-157 -- drawn, unused with w_cd = 0
        dev = src.device
        src_mean, tgt_mean = src.mean(dim=0, keepdim=True), tgt.to(dev).mean(dim=0, keepdim=True)
        ls = self.landmarks[0].to(dev).float() - src_mean
        lt = self.landmarks[1].to(dev).float() - tgt_mean
        if ls.shape != lt.shape or ls.dim() != 2 or ls.shape[1] != 3:
            raise ValueError("landmarks must be two [L, 3] arrays of matched points")
        n_src, nl = src.shape[0], ls.shape[0]
        body = torch.cat([src - src_mean, ls])
        src_in = torch.cat([body, -body.sum(dim=0, keepdim=True)]).contiguous()
        tgt_in = torch.cat([lt, -lt.sum(dim=0, keepdim=True)]).contiguous()
        sp = torch.arange(n_src, n_src + nl, dtype=torch.int32, device=dev)
        tp = torch.arange(nl, dtype=torch.int32, device=dev)
        flat = NDP.flat_parameters()
        solver = self._get_solver(1, src_in.shape[0], tgt_in.shape[0], profile_every=8 if timer else None,
                                  nn_mode=2, samples=max(int(config.samples), nl))
        prof0 = solver.profile() if timer else None
        if timer: timer.tic("ndp_fused")
        warped, iters, losses = solver.register([src_in], [tgt_in], [flat], [sp], [tp], src_samples=[nl], tgt_samples=[nl])
        if timer: timer.toc("ndp_fused")
        if timer: _feed_timer(timer, solver, prof0, int(iters.sum()))
        NDP.load_flat_parameters(flat)
        NDP.gradient_setup(optimized_level=-1)
        self.NDP, self.last_iters, self.last_losses = NDP, iters[0], losses[0]
        return warped[0][:n_src] + tgt_mean, {}, timer

    # ------------------------------------------------------------------------------------------
    def register_batch(self, pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]], seeds: Optional[Sequence[int]] = None,
                       host: bool = False):
        """Throughput entry point (not in the reference, which is strictly one pair at a time,
        eval_nolearned.py:70): registers independent pairs concurrently on one GPU.  With `seeds`,
        torch.manual_seed(seeds[p]) is applied before pair p's weights and permutations are drawn,
        which makes every pair's result independent of batching and of the rank it runs on.
        host=True: clouds are CPU tensors; host<->device copies happen inside the native call.
        Returns (list of warped clouds, iters [npairs, m], last loss [npairs, m])."""
        return self._run_prepared(self._prepare_batch(pairs, seeds, host), host)

    def _prepare_batch(self, pairs, seeds, host):
        """Host-side part of register_batch: per pair the pyramid's fresh weights (reference RNG order,
        nets.py:20-30,180-183) and the two sampling permutations (registration.py:156-157).  With per-pair seeds
        the pairs are independent random streams (a private torch.Generator seeded like torch.manual_seed would
        seed the global one: bit-identical draws), so they are prepared by a small thread pool; without seeds the
        global generator is consumed pair after pair, exactly like a sequential reference run."""
        config = self.config
        if config.w_reg > 0:
            raise NotImplementedError("register_batch covers the Chamfer-only NDP objective")
        dev = torch.device("cuda", self.device) if isinstance(self.device, int) else torch.device(self.device)
        n = len(pairs)
        per_pair = _params_per_pair(config)
        pin = host and torch.cuda.is_available()
        flat_all = torch.empty(n, per_pair, dtype=torch.float32, pin_memory=pin)      # one block: no stacking copy later

        def one(p):
            src, tgt = pairs[p]
            gen = None
            if seeds is not None:
                gen = torch.Generator()
                gen.manual_seed(int(seeds[p]))
            _init_flat_cpu(config, generator=gen, out=flat_all[p])
            sp = torch.randperm(src.shape[0], generator=gen)[:config.samples].to(torch.int32)
            tp = torch.randperm(tgt.shape[0], generator=gen)[:config.samples].to(torch.int32)
            return sp, tp

        if seeds is not None and n > 1:
            import os
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1, n)) as pool:
                perms = list(pool.map(one, range(n)))
        else:
            perms = [one(p) for p in range(n)]
        srcs, tgts, sps, tps = [], [], [], []
        for (src, tgt), (sp, tp) in zip(pairs, perms):
            if host:
                srcs.append(src.contiguous()); tgts.append(tgt.contiguous()); sps.append(sp); tps.append(tp)
            else:
                srcs.append(src.to(dev).contiguous()); tgts.append(tgt.to(dev).contiguous())
                sps.append(sp.to(dev)); tps.append(tp.to(dev))
        flats = flat_all if host else list(flat_all.to(dev))
        return srcs, tgts, flats, sps, tps

    def _run_prepared(self, prepared, host):
        srcs, tgts, flats, sps, tps = prepared
        dev = torch.device("cuda", self.device) if isinstance(self.device, int) else torch.device(self.device)
        self.src_pcd = srcs[0] if not host else srcs[0].to(dev)   # keeps _get_solver's device key valid
        solver = self._get_solver(len(srcs), max(s.shape[0] for s in srcs), max(t.shape[0] for t in tgts))
        warped, iters, losses = solver.register(srcs, tgts, flats, sps, tps, host=host, params_out=False)
        self.last_iters, self.last_losses = iters, losses
        return warped, iters, losses

    def register_batches(self, batches, seeds=None, host: bool = True):
        """Generator over a sequence (or iterator) of batches, each a list of (src, tgt) pairs: yields register_batch's
        result per batch.  The host-side work for batch k + 1 -- pulling it from the iterator (data loading), weight
        construction in the reference's RNG order, permutations -- runs in a worker thread while batch k is being
        optimised on the GPU (the native call releases the GIL).  Same results as calling register_batch per batch;
        `seeds` is a sequence / iterator of per-batch seed lists (or None).  The preparation uses torch's global CPU
        generator (like the reference): do not draw from it in the consuming thread while the generator is active."""
        from concurrent.futures import ThreadPoolExecutor
        it_b = iter(batches)
        it_s = iter(seeds) if seeds is not None else None

        def prepare_next():
            try:
                batch = next(it_b)
            except StopIteration:
                return None
            return self._prepare_batch(batch, next(it_s) if it_s is not None else None, host)

        with ThreadPoolExecutor(max_workers=1) as pool:
            nxt = pool.submit(prepare_next)
            while True:
                prepared = nxt.result()
                if prepared is None:
                    return
                nxt = pool.submit(prepare_next)
                yield self._run_prepared(prepared, host)

    # ------------------------------------------------------------------------------------------
    def _optimize_stepwise(self, timer=None):
        """registration.py:126-262 with landmarks and/or the nonrigidity regulariser: the
        reference's control flow over the CUDA autograd ops."""
        config = self.config
        max_break_count = config.max_break_count
        break_threshold_ratio = config.break_threshold_ratio
        NDP = self._build_pyramid()
        self.src_pcd = self.src_pcd.to(self.device)
        src_mean = self.src_pcd.mean(dim=0, keepdims=True)
        tgt_mean = self.tgt_pcd.mean(dim=0, keepdims=True)
        src_pcd = self.src_pcd - src_mean
        tgt_pcd = self.tgt_pcd - tgt_mean
        src = torch.randperm(src_pcd.shape[0])
        tgt = torch.randperm(tgt_pcd.shape[0])
        s_sample = src_pcd[src[: config.samples].to(src_pcd.device)].contiguous()
        t_sample = tgt_pcd[tgt[: config.samples].to(tgt_pcd.device)].contiguous()
        use_ldmk = self.landmarks is not None
        if use_ldmk:
            src_ldmk = (self.landmarks[0].to(src_pcd.device) - src_mean).contiguous()
            tgt_ldmk = (self.landmarks[1].to(src_pcd.device) - tgt_mean).contiguous()
        w_cd = _cfg_get(config, "w_cd", 0.0)
        iters_done, losses = [], []
        for level in range(NDP.n_hierarchy):
            NDP.gradient_setup(optimized_level=level)
            optimizer = optim.Adam(NDP.pyramid[level].parameters(), lr=config.lr)
            break_counter = 0
            loss_prev = 1e+6
            steps = 0
            for it in range(config.iters):
                if use_ldmk:
                    if w_cd > 0:
                        src_pts = torch.cat([src_ldmk, s_sample])
                        warped_pts, data = NDP.warp(src_pts, max_level=level, min_level=level)
                        warped_ldmk = warped_pts[: len(src_ldmk)]
                        s_sample_warped = warped_pts[len(src_ldmk):]
                        loss_ldmk = torch.mean(torch.sum((warped_ldmk - tgt_ldmk) ** 2, dim=-1))
                        loss_CD = compute_truncated_chamfer_distance(s_sample_warped[None].contiguous(),
                                                                     t_sample[None], trunc=config.trunc_cd)
                        loss = loss_ldmk + w_cd * loss_CD
                    else:
                        warped_ldmk, data = NDP.warp(src_ldmk, max_level=level, min_level=level)
                        loss = torch.mean(torch.sum((warped_ldmk - tgt_ldmk) ** 2, dim=-1))
                else:
                    if timer: timer.tic("lvl_warp")
                    s_sample_warped, data = NDP.warp(s_sample, max_level=level, min_level=level)
                    if timer: timer.toc("lvl_warp")
                    if timer: timer.tic("Chamfer")
                    loss = compute_truncated_chamfer_distance(s_sample_warped[None], t_sample[None], trunc=1e+9)
                    if timer: timer.toc("Chamfer")
                if level > 0 and config.w_reg > 0:
                    nonrigidity = data[level][1]
                    target = torch.zeros_like(nonrigidity)
                    reg_loss = BCE(nonrigidity, target)
                    loss = loss + config.w_reg * reg_loss
                # early stop (registration.py:225-232)
                lv = loss.item()
                if lv < 1e-4:
                    break
                if abs(loss_prev - lv) < loss_prev * break_threshold_ratio:
                    break_counter += 1
                if break_counter >= max_break_count:
                    break
                loss_prev = lv
                if timer: timer.tic("backprop")
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                if timer: timer.toc("backprop")
                steps += 1
            iters_done.append(steps)
            losses.append(lv if config.iters > 0 else float("nan"))
            if use_ldmk:
                src_ldmk = warped_ldmk.detach()
                if w_cd > 0:
                    s_sample = s_sample_warped.detach().contiguous()
            else:
                s_sample = s_sample_warped.detach()
        NDP.gradient_setup(optimized_level=-1)
        with torch.no_grad():
            warped_pcd, data = NDP.warp(src_pcd.contiguous())
        warped_pcd = warped_pcd + tgt_mean
        self.NDP = NDP
        self.last_iters = torch.tensor(iters_done, dtype=torch.int32)
        self.last_losses = torch.tensor(losses, dtype=torch.float32)
        iter_cnt = {}
        return warped_pcd, iter_cnt, timer


def _feed_timer(timer, solver, prof0, iterations: int) -> None:
    """Device time of the fused route under the reference's timer keys: `lvl_warp` (kernel 1), `Chamfer` (NN search +
    epilogue), `backprop` (backward + reduction + Adam), as registration.py:207-213, 234-238 tic/toc them once per
    iteration.  The solver brackets the kernels of every k-th iteration with CUDA events; the sampled means are
    scaled to the iterations executed, and the timers advance by that many calls (so avg() stays per iteration)."""
    (ms0, n0), (ms1, n1) = prof0, solver.profile()
    dn = n1 - n0
    if dn <= 0 or iterations <= 0:
        return
    per_iter = {k: (ms1[k] - ms0[k]) / dn * 1e-3 for k in ms1}            # seconds per iteration
    groups = {"lvl_warp": ("warp_fwd",), "Chamfer": ("nn_search", "chamfer_epilogue"), "backprop": ("warp_bwd", "reduce_adam")}
    for key, parts in groups.items():
        total = sum(per_iter[p] for p in parts) * iterations
        t = timer.timers[key] if hasattr(timer, "timers") else None
        if t is not None and all(hasattr(t, a) for a in ("total_time", "calls", "diff")):
            t.total_time += total; t.calls += iterations; t.diff = total / iterations
        else:
            timer.tictoc(key, total)


def _level_linears(config):
    W = config.width
    linears = [(W, 6)] + [(W, W)] * (config.depth - 1)
    if config.motion_type in ("Sim3", "SE3"):
        rdim = {"axis_angle": 3, "euler": 3, "quaternion": 4, "6D": 6}.get(config.rotation_format)
        if rdim is not None:
            linears.append((rdim, W))
        if config.motion_type == "Sim3":
            linears.append((1, W))
    linears.append((3, W))
    return linears


def _params_per_pair(config) -> int:
    return config.m * sum(o * i + o for o, i in _level_linears(config))


_INIT_PLANS = {}


def _init_plan(config):
    """Per configuration: where every final parameter of a level sits in the level's sequence of random draws, and
    its range.  The reference draws, per level, nn.Linear's default initialisation of every sub-module in construction
    order (weight U(+-1/sqrt(fan_in)) = kaiming_uniform(a=sqrt(5)), then bias U(+-1/sqrt(fan_in)); nets.py:75-101) and
    then Xavier uniform over every weight in parameters() order (nets.py:180-183): the weights are drawn twice, only the
    second draw survives."""
    import math
    linears = _level_linears(config)
    key = tuple(linears)
    plan = _INIT_PLANS.get(key)
    if plan is None:
        per_level = sum(o * i + o for o, i in linears)
        first, pos = [], 0
        for o, i in linears:
            first.append((pos, pos + o * i))              # (first weight draw, bias draw)
            pos += o * i + o
        idx = torch.empty(per_level, dtype=torch.long)
        a = torch.empty(per_level, dtype=torch.float32)
        p = 0
        for (o, i), (_, bpos) in zip(linears, first):
            idx[p:p + o * i] = torch.arange(pos, pos + o * i); pos += o * i                  # the Xavier draw
            a[p:p + o * i] = math.sqrt(3.0) * math.sqrt(2.0 / float(o + i)); p += o * i      # xavier_uniform_, gain 1
            idx[p:p + o] = torch.arange(bpos, bpos + o)
            a[p:p + o] = 1.0 / math.sqrt(i); p += o
        plan = _INIT_PLANS[key] = (per_level, pos, idx, 2.0 * a, -a)
    return plan


def _init_flat_cpu(config, generator=None, out=None) -> torch.Tensor:
    """Fresh weights of a whole pyramid drawn on the CPU in the reference's RNG order, flattened (level 0 first),
    without building nn.Modules.  ONE uniform_(0, 1) call draws the whole sequence (torch's CPU generator hands out one
    32-bit word per float32 sample whatever the tensor, so the stream is the reference's), then the surviving draws are
    gathered and mapped to their ranges with the same fused `x * (to - from) + from` torch's uniform_ evaluates: values
    and final generator state are bit-identical to `Deformation_Pyramid(...)` (tests/test_cabi_and_host.py compares the
    two bit for bit).  Three long kernels instead of 135 short ones per pair: the per-pair preparation threads of
    register_batch no longer queue behind the interpreter lock.  Consumes torch's global CPU generator -- or
    `generator`, a private stream."""
    per_level, seq_len, idx, scale, lo = _init_plan(config)
    flat = out if out is not None else torch.empty(config.m * per_level, dtype=torch.float32)
    u = torch.empty(config.m, seq_len, dtype=torch.float32).uniform_(0.0, 1.0, generator=generator)
    torch.addcmul(lo, torch.index_select(u, 1, idx), scale, out=flat.view(config.m, per_level))
    return flat


def _init_flat_modules(config) -> torch.Tensor:
    """The same through the nn.Module constructors (the reference's own path); kept as the cross-check."""
    from .nets import NDPLayer
    chunks = []
    for i in range(config.m):
        layer = NDPLayer(config.depth, config.width, config.k0, i + 1, config.rotation_format,
                         nonrigidity_est=False, motion=config.motion_type)
        chunks += [p.detach().reshape(-1) for p in layer.parameters()]
    return torch.cat(chunks).contiguous()
