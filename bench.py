#!/usr/bin/env python
"""Benchmark of the per-pair NDP hot path (BASELINE.json metric: registered point-cloud pairs/s for
8192-pt pairs, 9-level NDP.yaml pyramid with samples=8192, 500 iterations per level).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ndp|reference]

A "step" registers `--pairs` independent synthetic pairs per GPU (fixed-iteration mode A: early
stop disabled, exactly levels x iters Adam iterations per pair).  One JSON line on stdout (rank 0).
  value : pairs/s with the clouds, weights and permutations already resident in HBM
          (ndp_solver_register_device), device time from CUDA events, max over ranks.
  e2e   : pairs/s through Registration.register_batches(host=True): weight construction on the host,
          pinned-host -> device copies, the optimisation, device -> host read-back of the warped
          clouds, all inside the timed region (wall clock bracketed by synchronize, max over ranks);
          the host preparation of the next batch overlaps the GPU work of the current one.
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

import torch  # noqa: E402

METRIC = "registered point-cloud pairs/sec (8192-pt, 9-level NDP, 500 iters/level)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ndp", choices=["ndp", "reference"])
    ap.add_argument("--pairs", type=int, default=32, help="pairs registered concurrently per GPU per step")
    ap.add_argument("--points", type=int, default=8192)
    ap.add_argument("--levels", type=int, default=9)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--mode", default="fixed", choices=["fixed", "asconfigured"])
    ap.add_argument("--cpu-sample-iters", type=int, default=3, help="iterations per level of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nn-mode", type=int, default=0, help="0: exact culled NN search (default), 1: brute force")
    ap.add_argument("--mlp", default="tensor", choices=["tensor", "fp32"], help="tcgen05 tensor cores (default) or FP32 pipes")
    ap.add_argument("--tpc", type=int, default=None, help="override: tiles per backward CTA")
    ap.add_argument("--fwd-rounds", type=int, default=None, help="override: tile pairs per forward CTA")
    ap.add_argument("--streams", type=int, default=None, help="override: stream groups")
    return ap.parse_args()


def _profile(a):
    """ndp_solver_cfg execution profile selected from the batch size (identical in both arms' config)."""
    from deformationpyramid_b200.ops import execution_profile
    prof = execution_profile(a.pairs)
    for k, v in (("tiles_per_bwd_cta", a.tpc), ("fwd_rounds", a.fwd_rounds), ("streams", a.streams)):
        if v is not None:
            prof[k] = v
    return prof


def workload(a):
    return {"workload": f"synthetic {a.points}-pt src/tgt pairs (deformationpyramid_b200.synthetic), {a.levels}-level "
                        f"config/NDP.yaml pyramid with samples={a.points}, {a.iters} iters/level, "
                        + ("early stop disabled (fixed-iter mode A)" if a.mode == "fixed"
                           else "shipped early-stop thresholds (mode B)"),
            "points": a.points, "levels": a.levels, "iters_per_level": a.iters, "mode": a.mode,
            "pairs_per_step_per_gpu": a.pairs,
            "profile": _profile(a),
            "mlp": "tcgen05 fp16 hi/lo split, 3 partial products, fp32 accumulate (fp32-accurate)" if a.mlp == "tensor" else "fp32 pipes", "nn_search": "exact culled (Morton blocks + boxes + seeds)" if a.nn_mode == 0 else "brute force", "width": 128, "depth": 3, "motion": "SE3", "rotation": "axis_angle",
            "l2": "flushed (256 MiB write) between timed steps; one step streams >= 100 MiB of saved activations "
                  "and gradient partials per iteration (larger than L2)"}


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
def cpu_sample(a, knn_threads: int, iters_cap: int):
    """The oracle port (restatement of registration.py:126-262 on torch-CPU + the C kNN loop) on a
    bounded sample: `iters_cap` iterations per level of ONE pair of the benchmark workload."""
    from oracle import ndp_oracle as O
    from deformationpyramid_b200.synthetic import make_pair
    src, tgt = make_pair(0, a.points, a.points)
    cfg = O.NDPConfig(iters=a.iters, samples=a.points, m=a.levels,
                      max_break_count=10 ** 9 if a.mode == "fixed" else 15)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    res = O.optimize_pair(cfg, src, tgt, knn_threads=knn_threads, iters_cap=iters_cap)
    dt = time.perf_counter() - t0
    its = sum(len(c) for c in res.loss_curve)
    per_iter = dt / max(its, 1)
    pair_s = per_iter * a.levels * a.iters
    return {"seconds": dt, "iterations": its, "sec_per_iteration": per_iter, "pairs_per_s": 1.0 / pair_s}


def run_reference(a, rank):
    if rank != 0:
        return
    from oracle import ndp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    kt = min(cores, O.max_threads())
    for _ in range(a.warmup):
        cpu_sample(a, kt, 1)
    t, vals = 0.0, []
    for _ in range(a.steps):
        r = cpu_sample(a, kt, a.cpu_sample_iters)
        t += r["seconds"]; vals.append(r["pairs_per_s"])
    v = sum(vals) / len(vals)
    sample = (f"{a.levels} levels x {a.cpu_sample_iters} iterations of one {a.points}-pt pair per step (torch-CPU MLP/"
              f"autograd/Adam with {cores} threads + OpenMP C kNN with {kt} threads), per-iteration time linearly "
              f"extrapolated to {a.levels}x{a.iters} iterations")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": workload(a),
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    from deformationpyramid_b200 import ops
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
    mbc = 10 ** 9 if a.mode == "fixed" else 15
    cfg = ndp_config(samples=N, m=a.levels, iters=a.iters, max_break_count=mbc, device=local, **_profile(a))
    # pairs of this rank: global pair index = rank * B + p (independent units, no data-path collective)
    gids = [rank * B + p for p in range(B)]
    pairs = [make_pair(g, N, N) for g in gids]

    # ---- device-resident arm -------------------------------------------------------------------
    solver = ops.Solver(max_pairs=B, max_src_points=N, max_tgt_points=N, samples=N, levels=a.levels, k0=cfg.k0,
                        depth=cfg.depth, width=cfg.width, motion=cfg.motion_type, rotation_format=cfg.rotation_format,
                        iters=a.iters, max_break_count=mbc, break_threshold_ratio=cfg.break_threshold_ratio, lr=cfg.lr,
                        profile_every=16, nn_mode=a.nn_mode, mlp_mode=a.mlp, **_profile(a))
    d_src = [s.to(dev) for s, _ in pairs]
    d_tgt = [t.to(dev) for _, t in pairs]
    flats0, sps, tps = [], [], []
    for g in gids:
        torch.manual_seed(g)
        flats0.append(_init_flat_cpu(cfg).to(dev))
        sps.append(torch.randperm(N)[:N].to(torch.int32).to(dev))
        tps.append(torch.randperm(N)[:N].to(torch.int32).to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def dev_step():
        flats = [f.clone() for f in flats0]
        flush.fill_(1)                                   # L2 flush, outside the event bracket
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, its, last = solver.register(d_src, d_tgt, flats, sps, tps)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1), its, last

    for _ in range(a.warmup):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = solver.launch_count
    prof0, ns0 = solver.profile()
    ms_dev, its = 0.0, None
    for _ in range(a.steps):
        ms, its, last = dev_step()
        ms_dev += ms
    barrier()
    launches = solver.launch_count - l0
    prof1, ns1 = solver.profile()
    prof_pairs = solver.profiled_pairs
    clocks = sampler.stop() if rank == 0 else None
    solver.close()

    # ---- end-to-end arm: the reference-facing API with host buffers ------------------------------
    reg = Registration(cfg)
    h_pairs = [(s.pin_memory(), t.pin_memory()) for s, t in pairs]
    for _ in range(min(a.warmup, 1) if a.warmup else 0):
        reg.register_batch(h_pairs, seeds=gids, host=True)
    barrier()
    t0 = time.perf_counter()
    # every step = one batch through the public API: weights + permutations built on the host, pinned host ->
    # device copies, optimisation, device -> host read-back; the host preparation of step k + 1 overlaps the
    # GPU work of step k (Registration.register_batches), all of it inside the timed region
    for warped, _, last_e2e in reg.register_batches([h_pairs] * a.steps, seeds=[gids] * a.steps, host=True):
        if dist is not None:                              # the path's only collective: final metric gather
            g = [torch.empty_like(last_e2e[:, -1].to(dev)) for _ in range(world)]
            dist.all_gather(g, last_e2e[:, -1].contiguous().to(dev))
    barrier()
    s_e2e = time.perf_counter() - t0

    t_dev = torch.tensor([ms_dev / 1e3, s_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    sec_dev, sec_e2e = float(t_dev[0]), float(t_dev[1])
    total_pairs = B * world * a.steps
    value = total_pairs / sec_dev
    e2e = total_pairs / sec_e2e

    if rank == 0:
        pk, pk_src = peaks()
        P = 34694
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
        # dominant kernel: the tensor-core backward (largest share of the step, profiles/r01_launches_*.csv).
        # Algorithmic flops per point (kernel (3a)): dW and dH of the two 128x128 layers
        # 4 * 2*128*128, heads and their back-projection 2 * 2*6*128, input layer 2*128*6 = 135 680.
        bwd_flops = prof_pairs * N * (4 * 2 * 128 * 128 + 2 * 2 * 6 * 128 + 2 * 128 * 6)
        tensor_peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0))   # kernel timed inside a long step
        ach_tf = bwd_flops / max(t_bwd, 1e-12) / 1e12
        iso_ms = None
        if traffic.get("pairs_per_launch") == prof_pairs:
            ku = [(traffic.get(k) or {}).get("gpu__time_duration.sum") for k in ("ndp_warp_bwd_tc_kernel", "ndp_head_grad_kernel")]
            if all(v is not None for v in ku):
                iso_ms = sum(ku) * 1e-3
        roofline = {"bound": "tensor", "kernel": "ndp_head_grad_kernel + ndp_warp_bwd_tc_kernel (one backward launch)",
                    "achieved": ach_tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach_tf / tensor_peak,
                    "traffic": (traffic.get("ndp_warp_bwd_tc_kernel") or {}).get("dram_bytes_per_launch"),
                    "peak_source": pk_src + " bf16_tflops_sustained", "algorithmic_flops_per_launch": bwd_flops,
                    "pairs_per_launch": prof_pairs, "launch_ms": 1e3 * t_bwd,
                    "launch_ms_note": "sampled inside the step with the other stream groups' kernels interleaved on the same SMs; "
                                      "isolated = the ncu launch (profiles/kernel_traffic.json), same pairs per launch",
                    "isolated_launch_ms": iso_ms, "isolated_frac": (bwd_flops / (iso_ms * 1e-3) / 1e12 / tensor_peak) if iso_ms else None,
                    # the launches of the four stream groups overlap, so per-launch durations overstate the cost:
                    # the same ratio for the whole step = algorithmic MLP flops (forward 0.56 + backward 1.11 GFLOP per
                    # pair and iteration at N = 8192) of everything the step registered / the step's device time
                    "step_level": {"achieved": (B * iters_done * N / 8192.0 * 1.67e9) / (sec_dev / a.steps) / 1e12,
                                   "frac": (B * iters_done * N / 8192.0 * 1.67e9) / (sec_dev / a.steps) / 1e12 / tensor_peak,
                                   "unit": "TFLOP/s"},
                    "note": "fp32-accurate products are issued as 3 fp16 MMAs: issued tensor flops = 3x algorithmic"}
        # the metric's second half: one Chamfer call (NN search + epilogue) against the HBM roof
        alg_bytes = prof_pairs * (20 * (N + N) + 12 * N + 4)          # 425 988 B per pair at 8192^2
        achieved = alg_bytes / max(t_nn + t_ep, 1e-12) / 1e9
        evals = prof_pairs * 2.0 * N * N
        fp32_roof_evals = 148 * 128 * sm_clock * 1e6 / 8.0     # 8 FP32-pipe instructions per pair evaluation
        nn_names = ("ndp_nn_pruned_kernel", "ndp_chamfer_reduce_kernel") if a.nn_mode == 0 else ("ndp_nn_kernel", "ndp_chamfer_reduce_kernel")
        nn_traffic = [(traffic.get(k) or {}).get("dram_bytes_per_launch") for k in nn_names]
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": 1e3 * sec_dev / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(a),
               "clocks": clocks,
               "e2e": {"value": e2e, "unit": UNIT,
                       "h2d_bytes_per_step": B * (2 * N * 12 + 2 * N * 4 + a.levels * P * 4),
                       "d2h_bytes_per_step": B * (N * 12 + a.levels * P * 4 + a.levels * 8)},
               "gpu_launches": int(launches),
               "iterations_per_pair": iters_done,
               "roofline": roofline,
               "roofline_chamfer": {"bound": "hbm", "kernel": " + ".join(nn_names) + " (one Chamfer call)",
                                    "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": achieved / pk["hbm_gbs"],
                                    "traffic": sum(nn_traffic) if all(v is not None for v in nn_traffic) else None,
                                    "peak_source": pk_src, "algorithmic_bytes_per_launch": alg_bytes,
                                    "pairs_per_launch": prof_pairs, "launch_ms": 1e3 * (t_nn + t_ep),
                                    "note": "exact NN search is ALU-issue bound, not HBM bound (SURVEY.md 8d): see fp32"},
               "fp32": {"pair_evals_per_s": evals / max(t_nn, 1e-12), "roof_pair_evals_per_s": fp32_roof_evals,
                        "frac": evals / max(t_nn, 1e-12) / fp32_roof_evals, "sm_mhz_used": sm_clock,
                        "model": "brute-force equivalent pair evaluations (2NM per pair) / time, against 148 SM x 128 lanes x "
                                 "clock / 8 issue slots; the culled search skips most of them, so > 1 means 'faster than "
                                 "any brute-force kernel could be'"},
               "kernel_ms_per_launch": shares, "pairs_per_launch": prof_pairs}
        if not a.no_cpu_baseline:
            from oracle import ndp_oracle as O
            torch.set_num_threads(os.cpu_count() or 1)
            cores = os.cpu_count() or 1
            kt = min(cores, O.max_threads())
            cpu_sample(a, kt, 1)                              # warm the thread pools / the C library
            par = cpu_sample(a, kt, a.cpu_sample_iters)
            one = cpu_sample(a, 1, 1)
            out["cpu_baseline"] = {
                "value": par["pairs_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{a.levels} levels x {a.cpu_sample_iters} iterations of one {a.points}-pt pair "
                          f"({par['seconds']:.1f} s; torch-CPU MLP/autograd/Adam on {cores} threads, OpenMP C kNN on "
                          f"{kt} threads), per-iteration time extrapolated to {a.levels}x{a.iters} iterations",
                "single_thread_knn_pairs_per_s": one["pairs_per_s"],
                "single_thread_knn_note": "pytorch3d's CPU kNN is single-threaded: 9 levels x 1 iteration sample"}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
