"""CPU-only: the C-ABI library loads and exports every symbol include/ndp_b200.h declares (no compute
calls without a GPU), the host-side mirror constructs like the reference, config shim, loud failure
without a GPU."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from deformationpyramid_b200 import _lib, build as ndp_build
from deformationpyramid_b200.config import AttrDict, load_config, ndp_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "ndp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ndp_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_lib.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    path = ndp_build.build()
    lib = ctypes.CDLL(path)                      # loading needs no GPU
    for name in _header_symbols():
        assert hasattr(lib, name), name
    _lib.bind(lib)
    assert lib.ndp_version() >= 100
    # pure host arithmetic entry points may be called without a device
    cfg = _lib.LayerCfg(128, 3, 0, 0, 0, 2.0 ** -7, 0.001)
    assert lib.ndp_param_count(ctypes.byref(cfg)) == 34694
    cfg2 = _lib.LayerCfg(128, 3, 1, 1, 0, 2.0 ** -7, 0.001)
    assert lib.ndp_param_count(ctypes.byref(cfg2)) == 34823
    assert lib.ndp_saved_floats(ctypes.byref(cfg), 256) == 256 * 12     # depth 3 on the tensor cores: head vectors only (the backward recomputes)
    cfg5 = _lib.LayerCfg(128, 5, 0, 0, 0, 2.0 ** -7, 0.001)
    assert lib.ndp_saved_floats(ctypes.byref(cfg5), 256) == 2 * 5 * 16384 + 256 * 12    # other depths: fp16 hi/lo image sets (65536 B per tile and layer)
    bad = _lib.LayerCfg(64, 3, 0, 0, 0, 1.0, 0.001)
    assert lib.ndp_param_count(ctypes.byref(bad)) == -1
    assert b"width" in lib.ndp_last_error()


def test_dynamic_shared_memory_fits_the_opt_in_limit():
    """Every MLP kernel's dynamic shared memory fits sm_100's 227 KB per-block opt-in limit (cudaFuncSetAttribute fails
    with 'invalid argument' otherwise -- only visible on a GPU)."""
    lib = ctypes.CDLL(ndp_build.build())
    lib.ndp_debug_smem_bytes.restype = ctypes.c_longlong
    for which in range(6):
        n = lib.ndp_debug_smem_bytes(which)
        assert 0 < n <= 232448, (which, n)


def test_sass_shows_bulk_tma_and_fp32_pipeline():
    """The built cubin carries the Blackwell bulk-copy (UBLKCP) path and was built for sm_100a."""
    import subprocess
    path = ndp_build.build()
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out          # cp.async.bulk (TMA) staging of the MLP weights
    assert "FFMA" in out
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTCBAR"):      # tcgen05.mma / tcgen05.ld / tcgen05.st / tcgen05.commit
        assert mnemonic in out, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the GPU-less failure mode")
def test_product_path_fails_loudly_without_gpu():
    _lib._LIB = None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
    from deformationpyramid_b200.model.registration import Registration
    reg = Registration(ndp_config(device=torch.device("cpu"), samples=10, m=1, iters=1))
    reg.load_pcds(np.zeros((20, 3), np.float32), np.ones((20, 3), np.float32))
    with pytest.raises((RuntimeError, ValueError)):
        reg.register()


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "deformationpyramid_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/knn_oracle.c", "").replace("see oracle", "") or f == "build.py", \
                    os.path.join(dirpath, f)


def test_constructor_matches_reference_rng_stream(golden_dir):
    from deformationpyramid_b200.model.nets import NDPLayer
    G = np.load(os.path.join(golden_dir, "layers.npz"))
    for vi, meta in enumerate(G["meta"]):
        motion, fmt, nr, m, seed, depth = str(meta).split(",")
        torch.manual_seed(int(seed))
        layer = NDPLayer(int(depth), 128, -8, int(m), fmt, nonrigidity_est=bool(int(nr)), motion=motion)
        flat = torch.cat([p.detach().reshape(-1) for p in layer.parameters()]).numpy()
        assert hashlib.sha256(flat.tobytes()).hexdigest() == str(G[f"v{vi}_init_sha"]), meta
        assert [n for n, _ in layer.named_parameters()] == list(G[f"v{vi}_names"])


def test_fast_weight_init_equals_module_constructors_bit_for_bit():
    """register_batch draws a pyramid's weights without building nn.Modules; values AND the generator state
    afterwards must equal the Deformation_Pyramid constructor path (= the reference's RNG order)."""
    from deformationpyramid_b200.config import ndp_config
    from deformationpyramid_b200.model.registration import _init_flat_cpu, _init_flat_modules
    for kw in (dict(), dict(motion_type="Sim3", rotation_format="euler"), dict(rotation_format="6D"),
               dict(motion_type="sflow"), dict(depth=4, m=3), dict(rotation_format="quaternion", m=2)):
        cfg = ndp_config(samples=100, device=0, **kw)
        torch.manual_seed(11); a = _init_flat_cpu(cfg); ra = torch.rand(4)
        torch.manual_seed(11); b = _init_flat_modules(cfg); rb = torch.rand(4)
        assert torch.equal(a, b) and torch.equal(ra, rb), kw


def test_flatten_parameters_keeps_values_and_optimizer_semantics():
    from deformationpyramid_b200.model.nets import NDPLayer
    torch.manual_seed(0)
    layer = NDPLayer(3, 128, -8, 2, "axis_angle")
    before = [p.detach().clone() for p in layer.parameters()]
    flat = layer.flatten_parameters_()
    for p, b in zip(layer.parameters(), before):
        assert torch.equal(p, b)
    assert layer._flat_view(tuple(layer.parameters())) is flat
    opt = torch.optim.Adam(layer.parameters(), lr=0.1)
    for p in layer.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    assert layer._flat_view(tuple(layer.parameters())) is flat            # in-place update keeps the views
    assert torch.allclose(flat[:5], torch.cat([b.reshape(-1) for b in before])[:5] - 0.1, atol=1e-6)


def test_config_shim(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text("iters: 500\nlr: 0.01\nfolder: a\nexp_dir: !join [x, 3]\nsplit: {test: '4DMatch-F'}\n")
    c = load_config(str(p))
    assert c.iters == 500 and c.exp_dir == "x_3" and c.split.test == "4DMatch-F"
    c.device = 0
    assert c["device"] == 0
    with pytest.raises(AttributeError):
        c.missing
    d = ndp_config(samples=8192)
    assert d.samples == 8192 and d.m == 9 and d.k0 == -8 and d.motion_type == "SE3"


def test_install_as_model_aliases():
    import sys
    import deformationpyramid_b200 as pkg
    saved = {k: sys.modules.get(k) for k in ("model", "model.nets", "model.loss", "model.registration")}
    try:
        pkg.install_as_model()
        from model.nets import Deformation_Pyramid            # noqa: F401  (the reference scripts' imports)
        from model.loss import compute_truncated_chamfer_distance  # noqa: F401
        from model.registration import Registration           # noqa: F401
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_execution_profile_rule():
    """ops.execution_profile(pairs, samples): latency defaults below 24 pairs; from 24 pairs four stream groups and as many
    128-sample tiles per backward CTA (power of two, 1..16) as leaves ~64 CTAs per stream-group launch."""
    from deformationpyramid_b200.ops import execution_profile as e
    assert e(8, 8192) == dict(tiles_per_bwd_cta=0, fwd_rounds=0, streams=0)
    assert e(32, 8192) == dict(tiles_per_bwd_cta=8, fwd_rounds=4, streams=4)
    assert e(64, 8192) == dict(tiles_per_bwd_cta=16, fwd_rounds=8, streams=4)
    assert e(32, 2000) == dict(tiles_per_bwd_cta=2, fwd_rounds=1, streams=4)
    assert e(512, 8192)["tiles_per_bwd_cta"] == 16 and e(24, 200)["tiles_per_bwd_cta"] == 1
