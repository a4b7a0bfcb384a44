"""Drop-in alias: ``import torch_motion_correction`` resolves to the B200-native package.

Put ``<repo>/compat`` (and ``<repo>``) on ``PYTHONPATH`` to run code written against
teamtomo/torch-motion-correction unchanged.  Same ``__all__`` as the reference
(``src/torch_motion_correction/__init__.py:32-44``) plus the additive fused entry points."""

import sys as _sys

import torch_motion_correction_b200 as _impl
from torch_motion_correction_b200 import *  # noqa: F401,F403
from torch_motion_correction_b200 import (  # noqa: F401
    correct_motion,
    correct_motion_fast,
    correct_motion_slow,
    correct_motion_sum,
    correct_motion_two_grids,
    estimate_global_motion,
    estimate_local_motion,
    estimate_motion,
    estimate_motion_cross_correlation_patches,
    evaluate_deformation_field,
    get_pixel_shifts,
    motion_correct,
    read_deformation_field_from_csv,
    write_deformation_field_to_csv,
)

__version__ = _impl.__version__
__all__ = [
    "estimate_local_motion",
    "correct_motion",
    "correct_motion_two_grids",
    "correct_motion_fast",
    "correct_motion_slow",
    "get_pixel_shifts",
    "evaluate_deformation_field",
    "estimate_global_motion",
    "estimate_motion_cross_correlation_patches",
    "write_deformation_field_to_csv",
    "read_deformation_field_from_csv",
]

# the submodule paths the reference's tests import from
for _name in (
    "correct_motion", "estimate_motion_xc", "estimate_motion_optimizer", "deformation_field_utils", "utils",
    "patch_grid", "data_io", "optimization_state", "spline_grids",
):
    _mod = __import__(f"torch_motion_correction_b200.{_name}", fromlist=["_"])
    _sys.modules[f"{__name__}.{_name}"] = _mod
    # like the reference: the FUNCTION correct_motion shadows the submodule of the same name
    globals().setdefault(_name, _mod)
