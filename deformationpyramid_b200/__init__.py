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

from . import headless  # noqa: E402,F401  (no heavy imports)

__version__ = "0.1.0"


def install_as_model() -> None:
    """Make `from model.nets import ...`, `from model.loss import ...` and `from model.registration import ...`
    (eval_nolearned.py:8,11, shape_transfer.py:15-16) resolve to this package's mirror modules.

    When the reference checkout is on sys.path, ITS `model` package stays in place -- only the three sub-modules of
    the hot path are replaced, so `from model.geometry import *` (eval_nolearned.py:1), model.rigid_body etc. keep
    working.  Without the reference on sys.path the mirror package itself is aliased as `model`."""
    import importlib
    from . import model as mirror
    from .model import loss, nets, registration
    pkg = sys.modules.get("model")
    if pkg is None:
        try:
            pkg = importlib.import_module("model")          # the reference's (namespace) package, if importable
        except ImportError:
            pkg = None
    if pkg is None or pkg is mirror:
        pkg = mirror
        sys.modules["model"] = mirror
    for name, mod in (("nets", nets), ("loss", loss), ("registration", registration)):
        sys.modules["model." + name] = mod
        try:
            setattr(pkg, name, mod)
        except (AttributeError, TypeError):
            pass
