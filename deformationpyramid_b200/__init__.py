"""deformationpyramid_b200 -- B200 (sm_100a) implementation of the per-pair Neural Deformation
Pyramid hot path of rabbityl/DeformationPyramid behind the reference's own Python API.

    from deformationpyramid_b200.model.registration import Registration
    from deformationpyramid_b200.model.nets import Deformation_Pyramid
    from deformationpyramid_b200.model.loss import compute_truncated_chamfer_distance

or, to run the reference's eval_nolearned.py / shape_transfer.py unchanged, call
install_as_model() before they do `from model.nets import ...`.
Importing the package never loads the CUDA library; the first op does, and fails loudly if the
library or a CUDA device is missing (there is no CPU fallback).
"""
import sys

__version__ = "0.1.0"


def install_as_model() -> None:
    """Alias this package's model/ sub-package as the top-level `model` package the reference
    scripts import (eval_nolearned.py:8,11, shape_transfer.py:15-16)."""
    from . import model
    from .model import loss, nets, registration
    sys.modules["model"] = model
    sys.modules["model.nets"] = nets
    sys.modules["model.loss"] = loss
    sys.modules["model.registration"] = registration
