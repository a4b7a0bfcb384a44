"""Deformation-field utilities (mirror of the reference's ``deformation_field_utils.py``)."""

from __future__ import annotations

import torch

from . import _ops
from ._common import as_f32, grid_kind, resolve_device


def evaluate_deformation_field(deformation_field: torch.Tensor, tyx: torch.Tensor, grid_type: str = "catmull_rom") -> torch.Tensor:
    """(yx, nt, nh, nw) field evaluated at ``tyx (..., 3)`` in [0, 1] -> ``(..., 2)``.

    Reference: deformation_field_utils.py:9-39.
    """
    dev = resolve_device(deformation_field, None)
    field = as_f32(deformation_field, dev)
    return _ops.spline_eval(field, grid_kind(grid_type), as_f32(tyx, dev))


def evaluate_deformation_field_at_t(deformation_field: torch.Tensor, t: float, grid_shape, grid_type: str = "catmull_rom") -> torch.Tensor:
    """(2, h, w) lattice of shifts at one normalised time.  Reference: deformation_field_utils.py:42-93."""
    dev = resolve_device(deformation_field, None)
    field = as_f32(deformation_field, dev)
    h, w = grid_shape
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h, device=dev), torch.linspace(0, 1, w, device=dev), indexing="ij")
    tyx = torch.stack([torch.full_like(yy, float(t)), yy, xx], dim=-1)
    return _ops.spline_eval(field, grid_kind(grid_type), tyx).permute(2, 0, 1).contiguous()


def resample_deformation_field(deformation_field: torch.Tensor, target_resolution) -> torch.Tensor:
    """Catmull-Rom resampling to (nt, nh, nw) (always Catmull-Rom: reference quirk Q3).

    Reference: deformation_field_utils.py:96-126.
    """
    dev = resolve_device(deformation_field, None)
    field = as_f32(deformation_field, dev)
    nt, nh, nw = target_resolution
    tt, yy, xx = torch.meshgrid(
        torch.linspace(0, 1, nt, device=dev), torch.linspace(0, 1, nh, device=dev), torch.linspace(0, 1, nw, device=dev),
        indexing="ij",
    )
    out = _ops.spline_eval(field, 0, torch.stack([tt, yy, xx], dim=-1))
    return out.permute(3, 0, 1, 2).contiguous()


def image_shifts_to_deformation_field(shifts: torch.Tensor, pixel_spacing: float, device=None) -> torch.Tensor:
    """(t, 2) px -> (2, t, 1, 1) Angstrom.  Reference: deformation_field_utils.py:129-162."""
    if device is not None:
        shifts = shifts.to(device)
    return (shifts * pixel_spacing).transpose(0, 1)[:, :, None, None].contiguous()
