"""End-to-end drivers built from the reference-compatible entry points (additive API).

``estimate_motion`` = global (whole-frame XC) -> patch XC on the rigidly pre-corrected movie ->
optional spline-coefficient optimisation; ``motion_correct`` = estimate + fused warp-and-sum.
This is the workflow the reference sketches in ``examples/ttMotion.py:287-329,383-398``."""

from __future__ import annotations

import torch

from .correct_motion import correct_motion_sum
from .estimate_motion_xc import estimate_global_motion, estimate_motion_cross_correlation_patches


def estimate_motion(
    image: torch.Tensor,
    pixel_spacing: float,
    patch_sidelength: int = 1024,
    deformation_field_resolution: tuple[int, int, int] | None = None,
    n_iterations: int = 0,
    b_factor: float = 500,
    frequency_range: tuple[float, float] = (300, 10),
    grid_type: str = "bspline",
    optimizer_kwargs: dict | None = None,
    device: torch.device = None,
):
    """Returns ``(field (2, nt, nh, nw) Angstrom, patch_centres)``.

    With ``n_iterations > 0`` the patch-XC field initialises ``estimate_local_motion`` on a
    ``deformation_field_resolution`` spline grid (default (3, 5, 5), BASELINE config 2)."""
    global_field = estimate_global_motion(
        image, pixel_spacing, b_factor=b_factor, frequency_range=frequency_range, device=device
    )
    field, centres = estimate_motion_cross_correlation_patches(
        image, pixel_spacing, b_factor=b_factor, frequency_range=frequency_range, patch_sidelength=patch_sidelength,
        deformation_field=global_field, device=device,
    )
    if n_iterations > 0:
        from .estimate_motion_optimizer import estimate_local_motion

        resolution = deformation_field_resolution or (3, 5, 5)
        field = estimate_local_motion(
            image, pixel_spacing, (patch_sidelength, patch_sidelength), resolution, field, device=device,
            n_iterations=n_iterations, b_factor=b_factor, frequency_range=frequency_range, grid_type=grid_type,
            optimizer_kwargs=optimizer_kwargs,
        )
    return field, centres


def motion_correct(image: torch.Tensor, pixel_spacing: float, grid_type: str = "bspline", device=None, **estimate_kwargs):
    """Estimate + correct: returns ``(aligned frame sum (h, w), field)``."""
    field, _ = estimate_motion(image, pixel_spacing, grid_type=grid_type, device=device, **estimate_kwargs)
    total = correct_motion_sum(image, field, pixel_spacing, grid_type=grid_type, device=device)
    return total, field
