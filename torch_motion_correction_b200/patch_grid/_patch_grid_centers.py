"""Centres of a regular, evenly distributed grid of overlapping patches.

Reference: patch_grid/_patch_grid_centers.py:10-213 (``patch_grid_centers``, ``_patch_centers_1d``)."""

from __future__ import annotations

import torch


def patch_centers_1d(dim_length: int, patch_length: int, patch_step: int, distribute_patches: bool = True) -> torch.Tensor:
    """int64 centres along one axis: first at ``patch_length // 2``, then every ``patch_step``; the
    slack at the far end is spread over the patches (rounded linspace) when ``distribute_patches``."""
    first = patch_length // 2
    last = max(dim_length - first - 1, first)
    centers = torch.arange(first, last + 1, patch_step)
    if distribute_patches:
        slack = last - centers[-1]
        centers = centers + torch.round(torch.linspace(0, slack, steps=len(centers))).long()
    return centers


def patch_grid_centers(image_shape, patch_shape, patch_step, distribute_patches: bool = True, device=None) -> torch.Tensor:
    """(..., n_axes) int64 centres for a 2-D ``(h, w)`` or 3-D ``(d, h, w)`` grid."""
    if not (len(image_shape) == len(patch_shape) == len(patch_step)):
        raise ValueError("image shape, patch length and patch step are not the same length.")
    if len(image_shape) not in (2, 3):
        raise NotImplementedError("only 2D and 3D patches currently supported")
    axes = [
        patch_centers_1d(int(n), int(p), int(s), distribute_patches)
        for n, p, s in zip(image_shape, patch_shape, patch_step)
    ]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    return grid.to(device) if device is not None else grid
