"""B200-native drop-in for the hot path of ``torch_motion_correction``.

Same public names, argument meaning, tensor shapes and spline-grid outputs as the reference
package (``/root/reference/src/torch_motion_correction/__init__.py:12-44``); every kernel is
hand-written CUDA for sm_100a behind a C ABI (``include/tmc_b200.h``).  No CPU fallback.
"""

from .correct_motion import (
    correct_motion,
    correct_motion_fast,
    correct_motion_slow,
    correct_motion_sum,
    correct_motion_two_grids,
    get_pixel_shifts,
)
from .deformation_field_utils import (
    evaluate_deformation_field,
    evaluate_deformation_field_at_t,
    image_shifts_to_deformation_field,
    resample_deformation_field,
)

from .estimate_motion_xc import estimate_global_motion, estimate_motion_cross_correlation_patches
from .patch_grid import patch_grid_centers
from .pipeline import estimate_motion, motion_correct
from .utils import normalize_image

__version__ = "0.1.0"

__all__ = [
    "correct_motion",
    "correct_motion_two_grids",
    "correct_motion_fast",
    "correct_motion_slow",
    "correct_motion_sum",
    "get_pixel_shifts",
    "evaluate_deformation_field",
    "estimate_global_motion",
    "estimate_motion_cross_correlation_patches",
    "estimate_motion",
    "motion_correct",
]
