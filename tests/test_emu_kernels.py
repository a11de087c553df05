"""CPU-only checks of the kernel SOURCES (tiling, indexing, barriers, host orchestration) by running
them through the test-only CUDA-on-CPU shim (tests/cpu_emu) against the golden vectors of the
unmodified reference and against the oracle.  The same comparisons run on the real B200 through
the nvcc-built library in tests/test_gpu_parity.py; this file exists because the build container
has no GPU.  Nothing here is a product code path."""
import os

import numpy as np
import pytest
import torch

from deformationpyramid_b200 import ops
from deformationpyramid_b200.synthetic import make_pair
from oracle import ndp_oracle as O
from emu_util import emu_lib

from parity_cases import (check_layers_against_golden, check_layers_vs_oracle_depths, check_chamfer_against_golden, check_adam,
                          check_trajectory_teacher_forced, check_solver_against_oracle,
                          check_chamfer_vs_oracle_random, check_culled_search_equals_brute_force,
                          check_solver_repeatable, check_fp32_pipe_mode)


@pytest.fixture(scope="module")
def lib():
    return emu_lib()


def test_layers_golden(lib, golden_dir):
    check_layers_against_golden(lib, golden_dir, device="cpu")


def test_backward_accumulates_over_tiles(lib):
    """Gradients of several tiles accumulated in TMEM by one CTA (what 8192-point clouds use) on small inputs."""
    from deformationpyramid_b200 import ops
    try:
        ops.set_layer_tuning(4, 0, lib=lib)
        check_layers_vs_oracle_depths(lib, "cpu", cases=((3, 700), (2, 257)))
        ops.set_layer_tuning(2, 0, lib=lib)
        check_layers_vs_oracle_depths(lib, "cpu", cases=((3, 385),))
        # the throughput profile of large batches: 8 tiles per backward CTA, 2 tile-pair rounds per forward CTA
        ops.set_layer_tuning(8, 2, lib=lib)
        check_layers_vs_oracle_depths(lib, "cpu", cases=((3, 700),))
    finally:
        ops.set_layer_tuning(0, 0, lib=lib)


def test_layers_other_depths_vs_oracle(lib):
    check_layers_vs_oracle_depths(lib, "cpu", cases=((1, 130), (2, 257), (5, 385)))


def test_chamfer_golden(lib, golden_dir):
    check_chamfer_against_golden(lib, golden_dir, device="cpu")


def test_chamfer_random_vs_oracle(lib):
    check_chamfer_vs_oracle_random(lib, device="cpu", sizes=[(1, 1), (5, 700), (513, 129), (1100, 1030)])


def test_adam(lib):
    check_adam(lib, device="cpu")


def test_trajectory(lib, golden_dir):
    check_trajectory_teacher_forced(lib, golden_dir, device="cpu")


def test_config2_as_written_teacher_forced_sampled(lib, golden_dir):
    """Two of the 200 steps in the emulation (all 200 + the free-running solver on the GPU: tests/test_gpu_parity.py)."""
    from parity_cases import check_config2_teacher_forced
    check_config2_teacher_forced(lib, "cpu", golden_dir, stride=120, free_run=False)


def test_solver_small(lib):
    check_solver_against_oracle(lib, device="cpu", host=True, npairs=2, n=200, m=180, samples=130, levels=2,
                                iters=3, early_stop=False)


def test_solver_early_stop_and_ragged(lib):
    # samples > cloud size for pair 1 -> ragged counts; aggressive early stop -> ragged termination
    check_solver_against_oracle(lib, device="cpu", host=True, npairs=2, n=150, m=140, samples=160, levels=2,
                                iters=6, early_stop=True, ratio=0.05, max_break=2)


def test_solver_brute_force_mode(lib):
    check_solver_against_oracle(lib, device="cpu", host=True, npairs=1, n=200, m=180, samples=130, levels=2,
                                iters=3, early_stop=False, nn_mode=1)


def test_culled_search_equals_brute_force(lib):
    check_culled_search_equals_brute_force(lib, "cpu", n=230, m=200, samples=180, levels=2, iters=2)


def test_solver_nn_indices_bit_exact(lib):
    from parity_cases import check_solver_last_nn
    check_solver_last_nn(lib, "cpu", n=220, m=200, samples=192, levels=1, iters=2, dup=20)
    check_solver_last_nn(lib, "cpu", n=140, m=150, samples=128, levels=1, iters=2, lattice=True, dup=0)
    check_solver_last_nn(lib, "cpu", n=110, m=100, samples=96, levels=1, iters=1, nn_mode=1, dup=10)


def test_solver_repeatable_with_early_stop(lib):
    check_solver_repeatable(lib, "cpu", n=180, m=170, samples=128, levels=2, iters=4)      # the GPU suite runs the full-size variant


def test_fp32_pipe_mode(lib, golden_dir):
    check_fp32_pipe_mode(lib, "cpu", golden_dir)
