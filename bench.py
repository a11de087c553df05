#!/usr/bin/env python
"""Benchmark of the per-pair NDP hot path (BASELINE.json metric: registered point-cloud pairs/s for
8192-pt pairs, 9-level NDP.yaml pyramid with samples=8192, 500 iterations per level).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ndp|reference]

A "step" registers `--pairs` independent synthetic pairs per GPU (fixed-iteration mode A: early
stop disabled, exactly levels x iters Adam iterations per pair).  ONE JSON line on stdout (rank 0);
everything else any library prints (e.g. NCCL_DEBUG=INFO) goes to stderr.
  value : pairs/s with the clouds, weights and permutations already resident in HBM
          (ndp_solver_register_device), device time from CUDA events, max over ranks.
  e2e   : pairs/s through the public API with HOST buffers -- shard.evaluate over
          Registration.register_batches(host=True): pair sharding by rank, per-pair seeding, weight
          construction on the host, pinned-host -> device copies, the optimisation, device -> host
          read-back of the warped clouds, and the path's only collective (the final all_gather of the
          per-pair rows) all inside the timed region; the host work of batch k + 1 overlaps the GPU
          work of batch k.
  mode_b: the same workload with the shipped early-stop thresholds (config/NDP.yaml), iterations
          executed reported next to the CPU oracle's on the same pairs.
--impl reference times the CPU oracle port of the same path (oracle/ndp_oracle.py + the C kNN loop)
on the host cores over a bounded sample of the same workload (see cpu_baseline.sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly one JSON line: native libraries (NCCL's INFO log, when the caller asks for it) write to
# file descriptor 1 directly, so fd 1 is pointed at stderr for the life of the process and the line goes to a
# private duplicate of the original stdout
_JSON_FD = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)

import torch  # noqa: E402

METRIC = "registered point-cloud pairs/sec (8192-pt, 9-level NDP, 500 iters/level)"
P_LEVEL = 34694          # parameters of one SE3 / axis-angle level
UNIT = "pairs/s"


def emit(obj):
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ndp", choices=["ndp", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="pairs registered concurrently per GPU per step")
    ap.add_argument("--points", type=int, default=8192)
    ap.add_argument("--levels", type=int, default=9)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--mode", default="fixed", choices=["fixed", "asconfigured"])
    ap.add_argument("--cpu-sample-iters", type=int, default=50, help="iterations per level of the CPU sample (BASELINE.md 3.4)")
    ap.add_argument("--cpu-modeb-pairs", type=int, default=3, help="pairs the CPU oracle registers in full in mode B")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mode-b", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5-shaped leg (30 000-pt clouds, samples=2000, mode B)")
    ap.add_argument("--config5-batches", type=int, default=4)
    ap.add_argument("--nn-mode", type=int, default=0, help="0: exact culled NN search (default), 1: brute force")
    ap.add_argument("--mlp", default="tensor", choices=["tensor", "fp32"], help="tcgen05 tensor cores (default) or FP32 pipes")
    ap.add_argument("--tpc", type=int, default=None, help="override: tiles per backward CTA")
    ap.add_argument("--fwd-rounds", type=int, default=None, help="override: tile pairs per forward CTA")
    ap.add_argument("--streams", type=int, default=None, help="override: stream groups")
    return ap.parse_args()


def _profile(a):
    """ndp_solver_cfg execution profile selected from the batch size (identical in both arms' config)."""
    from deformationpyramid_b200.ops import execution_profile
    prof = execution_profile(a.pairs, a.points)
    for k, v in (("tiles_per_bwd_cta", a.tpc), ("fwd_rounds", a.fwd_rounds), ("streams", a.streams)):
        if v is not None:
            prof[k] = v
    return prof


def workload(a):
    """Identical in the ndp and the reference arm: nothing here depends on the environment or on which arm runs."""
    return {"workload": f"synthetic {a.points}-pt src/tgt pairs (deformationpyramid_b200.synthetic), {a.levels}-level "
                        f"config/NDP.yaml pyramid with samples={a.points}, {a.iters} iters/level, "
                        + ("early stop disabled (fixed-iter mode A)" if a.mode == "fixed"
                           else "shipped early-stop thresholds (mode B)"),
            "points": a.points, "levels": a.levels, "iters_per_level": a.iters, "mode": a.mode,
            "pairs_per_step_per_gpu": a.pairs, "profile": _profile(a),
            "mlp": "tcgen05 fp16 hi/lo split, 3 partial products, fp32 accumulate (fp32-accurate)" if a.mlp == "tensor" else "fp32 pipes",
            "nn_search": "exact culled (Morton blocks + boxes + seeds)" if a.nn_mode == 0 else "brute force",
            "width": 128, "depth": 3, "motion": "SE3", "rotation": "axis_angle",
            "l2": "flushed (256 MiB write) between timed steps: the per-iteration working set of a step (weights, "
                  f"Adam moments, gradient partials, clouds: ~{1.9 * a.pairs:.0f} MiB at {a.pairs} pairs) does not exceed the 126 MiB L2"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_sample(a, knn_threads: int, iters_cap, pair: int = 0, mode=None):
    """The oracle port (restatement of registration.py:126-262 on torch-CPU + the C kNN loop) on ONE pair of the
    benchmark workload (same synthetic pair, same per-pair seed as the GPU arm): `iters_cap` iterations per level
    (None: as the configuration runs, i.e. mode B in full)."""
    from oracle import ndp_oracle as O
    from deformationpyramid_b200.synthetic import make_pair
    mode = mode or a.mode
    src, tgt = make_pair(pair, a.points, a.points)
    cfg = O.NDPConfig(iters=a.iters, samples=a.points, m=a.levels,
                      max_break_count=10 ** 9 if mode == "fixed" else 15)
    torch.manual_seed(pair)
    timers = {}
    t0 = time.perf_counter()
    res = O.optimize_pair(cfg, src, tgt, knn_threads=knn_threads, iters_cap=iters_cap, timers=timers)
    dt = time.perf_counter() - t0
    its = sum(len(c) for c in res.loss_curve)
    per_iter = dt / max(its, 1)
    pair_s = per_iter * a.levels * a.iters if mode == "fixed" else dt
    return {"seconds": dt, "iterations": its, "adam_steps": int(sum(res.iters_per_level)), "sec_per_iteration": per_iter,
            "pairs_per_s": 1.0 / pair_s, "timers_s": {k: round(v, 4) for k, v in timers.items()}}


def cpu_mode_b(a, knn_threads: int, npairs: int):
    """Mode B (shipped early-stop thresholds) in full on the first `npairs` pairs of the workload."""
    runs = [cpu_sample(a, knn_threads, None, pair=p, mode="asconfigured") for p in range(npairs)]
    sec = sum(r["seconds"] for r in runs)
    return {"pairs": npairs, "value": npairs / sec, "unit": UNIT, "seconds": round(sec, 2),
            "adam_steps_per_pair": [r["adam_steps"] for r in runs],
            "timers_s": {k: round(sum(r["timers_s"].get(k, 0.0) for r in runs), 3) for k in ("lvl_warp", "Chamfer", "backprop")}}


def run_reference(a, rank):
    if rank != 0:
        return
    from oracle import ndp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    kt = min(cores, O.max_threads())
    for _ in range(a.warmup):
        cpu_sample(a, kt, 1)
    cap = a.cpu_sample_iters if a.mode == "fixed" else None
    t, vals, last = 0.0, [], None
    for _ in range(a.steps):
        last = cpu_sample(a, kt, cap)
        t += last["seconds"]; vals.append(last["pairs_per_s"])
    v = sum(vals) / len(vals)
    if a.mode == "fixed":
        sample = (f"{a.levels} levels x {a.cpu_sample_iters} iterations of one {a.points}-pt pair per step (torch-CPU MLP/"
                  f"autograd/Adam with {cores} threads + OpenMP C kNN with {kt} threads), per-iteration time linearly "
                  f"extrapolated to {a.levels}x{a.iters} iterations (BASELINE.md 3.4)")
    else:
        sample = f"one {a.points}-pt pair per step registered in full with the shipped early-stop thresholds ({last['adam_steps']} Adam steps)"
    emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
          "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic", "config": workload(a),
          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                           "spread": [min(vals), max(vals)], "timers_s_last_step": last["timers_s"]},
          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "gpu_launches": 0})


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    from deformationpyramid_b200 import ops, shard
    from deformationpyramid_b200.config import ndp_config
    from deformationpyramid_b200.model.registration import Registration, _init_flat_cpu
    from deformationpyramid_b200.synthetic import make_pair

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B, N = a.pairs, a.points
    prof = _profile(a)

    def make_cfg(mode):
        return ndp_config(samples=N, m=a.levels, iters=a.iters, max_break_count=10 ** 9 if mode == "fixed" else 15,
                          device=local, **prof)

    cfg = make_cfg(a.mode)
    # pairs of this rank: global pair index = rank * B + p (independent units, no data-path collective)
    gids = [rank * B + p for p in range(B)]
    pairs = [make_pair(g, N, N) for g in gids]

    def make_solver(mode, profile_every):
        c = make_cfg(mode)
        return ops.Solver(max_pairs=B, max_src_points=N, max_tgt_points=N, samples=N, levels=a.levels, k0=c.k0,
                          depth=c.depth, width=c.width, motion=c.motion_type, rotation_format=c.rotation_format,
                          iters=a.iters, max_break_count=c.max_break_count, break_threshold_ratio=c.break_threshold_ratio,
                          lr=c.lr, profile_every=profile_every, nn_mode=a.nn_mode, mlp_mode=a.mlp, **prof)

    # ---- device-resident arm -------------------------------------------------------------------
    solver = make_solver(a.mode, 16)
    d_src = [s.to(dev) for s, _ in pairs]
    d_tgt = [t.to(dev) for _, t in pairs]
    flats0, sps, tps = [], [], []
    for g in gids:
        torch.manual_seed(g)
        flats0.append(_init_flat_cpu(cfg).to(dev))
        sps.append(torch.randperm(N)[:N].to(torch.int32).to(dev))
        tps.append(torch.randperm(N)[:N].to(torch.int32).to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def dev_step(sv):
        flats = [f.clone() for f in flats0]
        flush.fill_(1)                                   # L2 flush, outside the event bracket
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, its, last = sv.register(d_src, d_tgt, flats, sps, tps)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), its, last

    for _ in range(a.warmup):
        dev_step(solver)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = solver.launch_count
    prof0, ns0 = solver.profile()
    nn0 = solver.nn_stats() if a.nn_mode == 0 else (0, 0, 0)
    ms_dev, its = 0.0, None
    for _ in range(a.steps):
        ms, its, last = dev_step(solver)
        ms_dev += ms
    barrier()
    launches = solver.launch_count - l0
    prof1, ns1 = solver.profile()
    nn1 = solver.nn_stats() if a.nn_mode == 0 else (0, 0, 0)
    prof_pairs = solver.profiled_pairs
    clocks = sampler.stop() if rank == 0 else None
    solver.close()

    # ---- the same kernels with the GPU to themselves: ONE stream group (every launch covers all B pairs and is the only
    #      work on the device), 24 iterations per level, every 4th iteration bracketed with CUDA events.  Not part of the
    #      timed region: these launch times are what a kernel costs when it does not share the SMs with three other
    #      groups' kernels (the in-step figures above include that sharing).
    solo = None
    if rank == 0 and a.mode == "fixed":
        c1 = make_cfg("fixed")
        prof_solo = dict(prof); prof_solo["streams"] = 1
        sv1 = ops.Solver(max_pairs=B, max_src_points=N, max_tgt_points=N, samples=N, levels=a.levels, k0=c1.k0, depth=c1.depth,
                         width=c1.width, motion=c1.motion_type, rotation_format=c1.rotation_format, iters=min(a.iters, 24),
                         max_break_count=10 ** 9, break_threshold_ratio=c1.break_threshold_ratio, lr=c1.lr, profile_every=4,
                         nn_mode=a.nn_mode, mlp_mode=a.mlp, **prof_solo)
        sv1.register(d_src, d_tgt, [f.clone() for f in flats0], sps, tps)
        p1, n1 = sv1.profile()
        solo = {k: v / max(n1, 1) for k, v in p1.items()}
        solo["pairs_per_launch"] = sv1.profiled_pairs
        sv1.close()

    # ---- end-to-end arm: the reference-facing API with host buffers, through the sharded evaluation loop ----------
    reg = Registration(cfg)
    all_pairs = {}                                        # global item index -> numpy clouds (this rank's only)

    def get_item(i):
        g = i % (B * world)                               # the step's pairs repeat; item i belongs to rank i % world
        if g not in all_pairs:
            s, t = make_pair(g, N, N)
            all_pairs[g] = dict(src_pcd=s.numpy(), tgt_pcd=t.numpy())
        return all_pairs[g]

    n_items = B * world * a.steps
    for i in shard.shard_indices(B * world, rank, world):
        get_item(i)
    if a.warmup:
        shard.evaluate(reg, B * world, get_item, rank=rank, world=world, batch=B, base_seed=0, compute_metrics=False,
                       gather_device=dev, host=True)
    barrier()
    t0 = time.perf_counter()
    rows, _ = shard.evaluate(reg, n_items, get_item, rank=rank, world=world, batch=B, base_seed=0, compute_metrics=False,
                             gather_device=dev, host=True, checksum=True)
    barrier()
    s_e2e = time.perf_counter() - t0
    assert rows.shape[0] == n_items

    # SURVEY.md section 4, T4 on hardware: a pair's result is independent of the rank / batch position it ran in.
    # Rank 0 re-registers the batch that holds the LAST rank's first pair of the first step (same seeds, same batch
    # size => same execution profile) in a different order and compares the warped cloud's checksum bit for bit.
    t4 = None
    if rank == 0:
        foreign = [i for i in range(B * world) if i % world == world - 1][:B]
        while len(foreign) < B:
            foreign.append(foreign[-1])
        order = foreign[::-1]
        hp = []
        for i in order:
            s, t = make_pair(i % (B * world), N, N)
            hp.append((s.pin_memory(), t.pin_memory()))
        w2, _, _ = reg.register_batch(hp, seeds=[0 + i for i in order], host=True)
        want = float(rows[foreign[0], -1])
        got = shard.cloud_checksum(w2[len(order) - 1])
        t4 = {"pair": foreign[0], "owner_rank": world - 1, "bit_identical": bool(got == want)}
        assert t4["bit_identical"], ("pair result depends on the rank / batch position", t4, got, want)

    t_dev = torch.tensor([ms_dev / 1e3, s_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    sec_dev, sec_e2e = float(t_dev[0]), float(t_dev[1])
    total_pairs = B * world * a.steps
    value = total_pairs / sec_dev
    e2e = total_pairs / sec_e2e

    # ---- mode B (shipped early-stop thresholds): one step, device-resident and end to end --------------------
    mode_b = None
    if a.mode == "fixed" and not a.no_mode_b:
        sb = make_solver("asconfigured", 0)
        dev_step(sb)
        barrier()
        ms_b, its_b, _ = dev_step(sb)
        barrier()
        sb.close()
        regb = Registration(make_cfg("asconfigured"))
        shard.evaluate(regb, B * world, get_item, rank=rank, world=world, batch=B, base_seed=0, compute_metrics=False,
                       gather_device=dev, host=True)
        barrier()
        tb0 = time.perf_counter()
        shard.evaluate(regb, 4 * B * world, get_item, rank=rank, world=world, batch=B, base_seed=0, compute_metrics=False,
                       gather_device=dev, host=True)
        barrier()
        tb = torch.tensor([ms_b / 1e3, time.perf_counter() - tb0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        mode_b = {"value": B * world / float(tb[0]), "e2e": 4 * B * world / float(tb[1]), "unit": UNIT,
                  "adam_steps_per_pair_mean": float(its_b.sum()) / B,
                  "adam_steps_first_pairs": [int(v) for v in its_b.sum(dim=1)[:a.cpu_modeb_pairs]],
                  "note": "same pairs, weights and permutations as mode A; shipped thresholds max_break_count=15, "
                          "break_threshold_ratio=0.001 (config/NDP.yaml:10-11); e2e over four batches per rank"}

    # ---- config-5-shaped leg (BASELINE.json configs[4]; eval_nolearned.py:70-143 with config/NDP.yaml as shipped): clouds of
    #      30 000 points (the cap of _4dmatch.py:30), samples = 2000 (15.6 ragged tiles), 9 levels, <= 500 iterations per level
    #      with the shipped early-stop thresholds, through shard.evaluate with host buffers; the 4DMatch data is not
    #      available offline, the clouds are synthetic ------------------------------------------------------------------
    config5 = None
    if a.mode == "fixed" and not a.no_config5:
        NP5, B5 = 30000, B
        prof5 = {k: v for k, v in (("tiles_per_bwd_cta", a.tpc), ("fwd_rounds", a.fwd_rounds), ("streams", a.streams)) if v is not None}
        cfg5 = ndp_config(device=local, **prof5)                     # NDP.yaml defaults: samples 2000, m 9, iters 500, early stop on; the
        # execution profile follows Registration's rule for (pairs, samples) unless overridden on the command line
        reg5 = Registration(cfg5)
        pairs5 = {}

        def get_item5(i):
            g = i % (B5 * world)
            if g not in pairs5:
                s, t = make_pair(5000 + g, NP5, NP5)
                pairs5[g] = dict(src_pcd=s.numpy(), tgt_pcd=t.numpy())
            return pairs5[g]

        for i in shard.shard_indices(B5 * world, rank, world):
            get_item5(i)
        shard.evaluate(reg5, B5 * world, get_item5, rank=rank, world=world, batch=B5, base_seed=0, compute_metrics=False,
                       gather_device=dev, host=True)
        barrier()
        t50 = time.perf_counter()
        nb5 = a.config5_batches
        shard.evaluate(reg5, nb5 * B5 * world, get_item5, rank=rank, world=world, batch=B5, base_seed=0, compute_metrics=False,
                       gather_device=dev, host=True)
        barrier()
        t5 = torch.tensor([time.perf_counter() - t50], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        config5 = {"workload": f"synthetic {NP5}-pt clouds, config/NDP.yaml as shipped (samples=2000, m=9, iters<=500, max_break_count=15), "
                               f"{B5} pairs per batch per GPU, {nb5} batches per rank",
                   "e2e": {"value": nb5 * B5 * world / float(t5[0]), "unit": UNIT,
                           "h2d_bytes_per_batch": B5 * (2 * NP5 * 12 + 2 * 2000 * 4 + a.levels * P_LEVEL * 4),
                           "d2h_bytes_per_batch": B5 * (NP5 * 12 + a.levels * 8)},
                   "adam_steps_per_pair_mean": float(reg5.last_iters.sum()) / B5,
                   "api": "shard.evaluate(Registration, host=True)"}

    if rank == 0:
        pk, pk_src = peaks()
        P = P_LEVEL
        iters_done = int(its.sum()) // B if its is not None else a.levels * a.iters
        n_s = max(ns1 - ns0, 1)
        shares = {k: (prof1[k] - prof0[k]) / n_s for k in prof1}      # ms per sampled launch (prof_pairs pairs each)
        t_nn, t_ep, t_bwd = shares["nn_search"] * 1e-3, shares["chamfer_epilogue"] * 1e-3, shares["warp_bwd"] * 1e-3
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        traffic = {}
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f)
        sm_clock = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
        bwd_kernel = "ndp_warp_bwd_rc_kernel" if a.mlp == "tensor" else "ndp_warp_bwd_kernel"
        # dominant kernel: the tensor-core backward (largest share of the step, profiles/*launches*.csv).
        # Algorithmic flops per point (kernel (3a)): dW and dH of the two 128x128 layers
        # 4 * 2*128*128, heads and their back-projection 2 * 2*6*128, input layer 2*128*6 = 135 680.  (The kernel
        # REBUILDS the activations, +65 536 flop per point, and issues every product as 3 fp16 MMAs: neither counts.)
        bwd_flops = prof_pairs * N * (4 * 2 * 128 * 128 + 2 * 2 * 6 * 128 + 2 * 128 * 6)
        tensor_peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))   # kernel timed inside a long step
        ach_tf = bwd_flops / max(t_bwd, 1e-12) / 1e12
        iso_ms = None
        if traffic.get("pairs_per_launch") == prof_pairs:
            ku = [(traffic.get(k) or {}).get("gpu__time_duration.sum") for k in (bwd_kernel, "ndp_head_grad_kernel")]
            if all(v is not None for v in ku):
                iso_ms = sum(ku) * 1e-3
        step_flops = B * iters_done * N / 8192.0 * 1.67e9
        roofline = {"bound": "tensor", "kernel": f"ndp_head_grad_kernel + {bwd_kernel} (one backward launch)",
                    "achieved": ach_tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach_tf / tensor_peak,
                    "traffic": (traffic.get(bwd_kernel) or {}).get("dram_bytes_per_launch"),
                    "peak_source": pk_src + " bf16_tflops_sustained", "algorithmic_flops_per_launch": bwd_flops,
                    "algorithmic_bytes_per_launch": prof_pairs * (24 * N + 6 * 138776 + (64 // max(prof.get("tiles_per_bwd_cta") or 4, 1)) * 138776),
                    "pairs_per_launch": prof_pairs, "launch_ms": 1e3 * t_bwd,
                    "launch_ms_note": "sampled inside the step with the other stream groups' kernels interleaved on the same SMs; "
                                      "isolated = the ncu launch (profiles/kernel_traffic.json), same pairs per launch",
                    "isolated_launch_ms": iso_ms, "isolated_frac": (bwd_flops / (iso_ms * 1e-3) / 1e12 / tensor_peak) if iso_ms else None,
                    # the same launch covering ALL pairs of the step on a single stream (nothing else on the device), measured
                    # live with CUDA events in this run, outside the timed region
                    "solo": ({"pairs_per_launch": solo["pairs_per_launch"], "launch_ms": solo["warp_bwd"],
                              "achieved": solo["pairs_per_launch"] * N * 135680 / (solo["warp_bwd"] * 1e-3) / 1e12,
                              "frac": solo["pairs_per_launch"] * N * 135680 / (solo["warp_bwd"] * 1e-3) / 1e12 / tensor_peak,
                              "kernel_ms_per_launch": {k: v for k, v in solo.items() if k != "pairs_per_launch"}}
                             if solo else None),
                    # the launches of the stream groups overlap, so per-launch durations overstate the cost: the same
                    # ratio for the whole step = algorithmic MLP flops (forward 0.56 + backward 1.11 GFLOP per pair and
                    # iteration at N = 8192) of everything the step registered / the step's device time
                    "step_level": {"achieved": step_flops / (sec_dev / a.steps) / 1e12,
                                   "frac": step_flops / (sec_dev / a.steps) / 1e12 / tensor_peak, "unit": "TFLOP/s"},
                    "note": "fp32-accurate products are issued as 3 fp16 MMAs and the activations are recomputed: issued tensor "
                            "flops = 4.4x algorithmic; SS-mode operand fetch (shared-memory bandwidth), not the tensor pipe, binds the kernel (DESIGN.md)"}
        # the metric's second half: one Chamfer call (NN search + epilogue) against the HBM roof
        alg_bytes = prof_pairs * (20 * (N + N) + 12 * N + 4)          # 425 988 B per pair at 8192^2
        achieved = alg_bytes / max(t_nn + t_ep, 1e-12) / 1e9
        nn_names = ("ndp_nn_pruned_kernel", "ndp_chamfer_reduce_kernel") if a.nn_mode == 0 else ("ndp_nn_kernel", "ndp_chamfer_reduce_kernel")
        nn_traffic = [(traffic.get(k) or {}).get("dram_bytes_per_launch") for k in nn_names]
        nn_work = None
        if a.nn_mode == 0 and nn1[1] > nn0[1]:
            evals, qblocks = nn1[0] - nn0[0], nn1[1] - nn0[1]
            searches = qblocks / (2.0 * (N // 32))                   # (pair, iteration) searches behind the counters
            per_search_s = t_nn / max(prof_pairs, 1)                 # sampled launch time / pairs per launch
            ipe = 4.5                                                # instructions per evaluation (csrc/ndp_spatial.cu: packed FP32)
            issue_roof = 148 * 128 * sm_clock * 1e6 / ipe            # lane-instructions per second / instructions per evaluation
            lds_roof = 148 * sm_clock * 1e6 * 32 / 3.0               # shared-memory pipe: 3 cycles per candidate broadcast to 32 lanes
            rate = evals / searches / max(per_search_s, 1e-12)
            nn_work = {"candidates_per_query": evals / (qblocks * 32.0), "brute_force_candidates_per_query": N,
                       "instructions_per_eval": ipe, "mean_blocks_scanned_per_warp": evals / 1024.0 / max(qblocks, 1),
                       "max_blocks_scanned_by_one_warp": nn1[2],
                       "pair_evals_issued_per_s": rate,
                       "issue_roof_pair_evals_per_s": issue_roof, "issue_frac": rate / issue_roof,
                       "smem_pipe_roof_pair_evals_per_s": lds_roof, "smem_pipe_frac": rate / lds_roof,
                       "model": "distance evaluations the culled search actually issues (32 x 32 per scanned block) per second of its "
                                "launch time, against (a) 148 SM x 128 lanes x clock / 4.5 instructions per evaluation and (b) the "
                                "shared-memory pipe that broadcasts the candidates (12 bytes per candidate to each of 32 lanes = 3 "
                                "cycles); an isolated launch is bound by neither: its slowest warp walks max_blocks blocks one after "
                                "the other (DESIGN.md)"}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": 1e3 * sec_dev / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(a),
               "clocks": clocks,
               "e2e": {"value": e2e, "unit": UNIT,
                       "h2d_bytes_per_step": B * (2 * N * 12 + 2 * N * 4 + a.levels * P * 4),
                       "d2h_bytes_per_step": B * (N * 12 + a.levels * P * 4 + a.levels * 8),
                       "api": "shard.evaluate(Registration, host=True): pinned host clouds in, warped clouds out, final all_gather of the rows"},
               "gpu_launches": int(launches),
               "iterations_per_pair": iters_done,
               "roofline": roofline,
               "roofline_chamfer": {"bound": "hbm", "kernel": " + ".join(nn_names) + " (one Chamfer call)",
                                    "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": achieved / pk["hbm_gbs"],
                                    "traffic": sum(nn_traffic) if all(v is not None for v in nn_traffic) else None,
                                    "peak_source": pk_src, "algorithmic_bytes_per_launch": alg_bytes,
                                    "pairs_per_launch": prof_pairs, "launch_ms": 1e3 * (t_nn + t_ep),
                                    "note": "exact NN search is ALU-issue bound, not HBM bound (SURVEY.md 8d): see nn_search_work"},
               "nn_search_work": nn_work,
               "kernel_ms_per_launch": shares, "pairs_per_launch": prof_pairs,
               "rank_independence": t4, "mode_b": mode_b, "config5": config5}
        if not a.no_cpu_baseline and world == 1:
            from oracle import ndp_oracle as O
            torch.set_num_threads(os.cpu_count() or 1)
            cores = os.cpu_count() or 1
            kt = min(cores, O.max_threads())
            cpu_sample(a, kt, 1)                              # warm the thread pools / the C library
            par = cpu_sample(a, kt, a.cpu_sample_iters if a.mode == "fixed" else None)
            one = cpu_sample(a, 1, 1)
            out["cpu_baseline"] = {
                "value": par["pairs_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{a.levels} levels x {a.cpu_sample_iters} iterations of one {a.points}-pt pair "
                          f"({par['seconds']:.1f} s; torch-CPU MLP/autograd/Adam on {cores} threads, OpenMP C kNN on "
                          f"{kt} threads), per-iteration time extrapolated to {a.levels}x{a.iters} iterations (BASELINE.md 3.4)",
                "timers_s": par["timers_s"],
                "single_thread_knn_pairs_per_s": one["pairs_per_s"],
                "single_thread_knn_note": "pytorch3d's CPU kNN is single-threaded: 9 levels x 1 iteration sample"}
            if mode_b is not None and a.cpu_modeb_pairs > 0:
                cb = cpu_mode_b(a, kt, a.cpu_modeb_pairs)
                out["cpu_baseline"]["mode_b"] = cb
                mode_b["adam_steps_cpu_oracle_first_pairs"] = cb["adam_steps_per_pair"]
        emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
