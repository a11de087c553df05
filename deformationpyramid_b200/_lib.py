"""ctypes binding of libndp_b200.so (include/ndp_b200.h).

The product path has NO CPU fallback: load() raises if the nvcc-built library is missing or no
CUDA device is visible.  bind() only declares prototypes; tests reuse it for the CPU-emulated
build of the same sources (tests/cpu_emu), which the product never loads.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libndp_b200.so")

MOTION = {"SE3": 0, "Sim3": 1, "sflow": 2}
ROT_FORMAT = {"axis_angle": 0, "euler": 1, "quaternion": 2, "6D": 3}

EXPORTS = [
    "ndp_last_error", "ndp_version", "ndp_param_count", "ndp_pack_count",
    "ndp_saved_floats", "ndp_set_mlp_mode", "ndp_get_mlp_mode", "ndp_set_layer_tuning", "ndp_backward_workspace_bytes", "ndp_chamfer_workspace_bytes",
    "ndp_pack_params", "ndp_layer_forward", "ndp_layer_backward", "ndp_chamfer", "ndp_adam_step",
    "ndp_solver_create", "ndp_solver_destroy", "ndp_solver_params_per_pair",
    "ndp_solver_register_host", "ndp_solver_register_device", "ndp_solver_last_nn", "ndp_solver_losses",
    "ndp_solver_launch_count", "ndp_solver_profile", "ndp_solver_profiled_pairs", "ndp_solver_nn_stats",
]


class LayerCfg(ctypes.Structure):
    _fields_ = [("width", c_int32), ("depth", c_int32), ("motion", c_int32), ("rot_format", c_int32),
                ("nonrigidity", c_int32), ("freq", c_float), ("mlp_scale", c_float)]


class SolverCfg(ctypes.Structure):
    _fields_ = [("max_pairs", c_int32), ("max_src_points", c_int32), ("max_tgt_points", c_int32),
                ("samples", c_int32), ("levels", c_int32), ("k0", c_int32), ("depth", c_int32),
                ("width", c_int32), ("motion", c_int32), ("rot_format", c_int32), ("iters", c_int32),
                ("max_break_count", c_int32), ("break_threshold_ratio", c_float), ("lr", c_double),
                ("trunc", c_float), ("record_loss", c_int32), ("profile_every", c_int32), ("nn_mode", c_int32),
                ("mlp_mode", c_int32), ("tiles_per_bwd_cta", c_int32), ("fwd_rounds", c_int32), ("streams", c_int32)]


def bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    P = POINTER
    lib.ndp_last_error.restype = c_char_p
    lib.ndp_last_error.argtypes = []
    lib.ndp_version.restype = c_int32
    for name in ("ndp_param_count", "ndp_pack_count"):
        getattr(lib, name).restype = c_int64
        getattr(lib, name).argtypes = [P(LayerCfg)]
    lib.ndp_saved_floats.restype = c_int64
    lib.ndp_saved_floats.argtypes = [P(LayerCfg), c_int64]
    lib.ndp_set_mlp_mode.argtypes = [c_int32]
    lib.ndp_set_mlp_mode.restype = ctypes.c_int
    lib.ndp_get_mlp_mode.restype = c_int32
    lib.ndp_set_layer_tuning.argtypes = [c_int32, c_int32]
    lib.ndp_set_layer_tuning.restype = ctypes.c_int
    lib.ndp_backward_workspace_bytes.restype = c_int64
    lib.ndp_backward_workspace_bytes.argtypes = [P(LayerCfg), c_int64]
    lib.ndp_chamfer_workspace_bytes.restype = c_int64
    lib.ndp_chamfer_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.ndp_pack_params.argtypes = [P(LayerCfg), c_void_p, c_void_p, c_void_p]
    lib.ndp_layer_forward.argtypes = [P(LayerCfg), c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_void_p, c_void_p]
    lib.ndp_layer_backward.argtypes = [P(LayerCfg), c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ndp_chamfer.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_float, c_float, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ndp_adam_step.argtypes = [P(LayerCfg), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                  c_double, c_double, c_double, c_double, c_void_p, c_void_p]
    lib.ndp_solver_create.argtypes = [P(SolverCfg), P(c_void_p)]
    lib.ndp_solver_destroy.argtypes = [c_void_p]
    lib.ndp_solver_destroy.restype = None
    lib.ndp_solver_params_per_pair.argtypes = [c_void_p]
    lib.ndp_solver_params_per_pair.restype = c_int64
    lib.ndp_solver_launch_count.argtypes = [c_void_p]
    lib.ndp_solver_launch_count.restype = c_int64
    pp = P(c_void_p)
    lib.ndp_solver_register_host.argtypes = [c_void_p, c_int32, pp, P(c_int32), pp, P(c_int32), pp, pp,
                                             P(c_int32), P(c_int32), c_void_p, c_int32, pp, c_void_p, c_void_p, c_void_p]
    lib.ndp_solver_register_device.argtypes = [c_void_p, c_int32, pp, P(c_int32), pp, P(c_int32), pp, pp,
                                               P(c_int32), P(c_int32), pp, pp, c_void_p, c_void_p, c_void_p]
    lib.ndp_solver_last_nn.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ndp_solver_losses.argtypes = [c_void_p, c_int32, c_void_p, c_void_p]
    lib.ndp_solver_profile.argtypes = [c_void_p, P(c_double), P(c_int64)]
    lib.ndp_solver_profile.restype = ctypes.c_int
    lib.ndp_solver_nn_stats.argtypes = [c_void_p, P(c_int64), P(c_int64), P(c_int64)]
    lib.ndp_solver_nn_stats.restype = ctypes.c_int
    lib.ndp_solver_profiled_pairs.argtypes = [c_void_p]
    lib.ndp_solver_profiled_pairs.restype = c_int32
    for name in ("ndp_pack_params", "ndp_layer_forward", "ndp_layer_backward", "ndp_chamfer",
                 "ndp_adam_step", "ndp_solver_create", "ndp_solver_register_host",
                 "ndp_solver_register_device", "ndp_solver_losses", "ndp_solver_last_nn"):
        getattr(lib, name).restype = ctypes.c_int
    return lib


_LIB = None


def load() -> ctypes.CDLL:
    """The product library.  Fails loudly -- there is no other compute path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m deformationpyramid_b200.build` "
            "(nvcc, sm_100a).  deformationpyramid_b200 has no CPU fallback.")
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("deformationpyramid_b200 needs a CUDA device (B200, sm_100a); none is visible "
                           "and there is no CPU fallback.")
    lib = bind(ctypes.CDLL(LIB_PATH))
    lib._ndp_requires_cuda = True
    _LIB = lib
    return lib


def check(lib, rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.ndp_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")
