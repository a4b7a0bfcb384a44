"""Motion correction with a deformation field (mirror of the reference's ``correct_motion.py``).

All arithmetic runs in hand-written sm_100a kernels behind the C ABI (``csrc/warp.cu``,
``csrc/spline.cu``, ``csrc/fourier.cu``); this file only validates arguments and owns tensors.
"""

from __future__ import annotations

import torch

from . import _ops
from ._common import as_f32, grid_kind, resolve_device


def _movie(image: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if image.ndim != 3:
        raise ValueError(f"image must be (t, h, w), got {tuple(image.shape)}")
    return as_f32(image, dev)


def _field(deformation_grid: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if deformation_grid.ndim != 4 or deformation_grid.shape[0] != 2:
        raise ValueError(f"deformation_grid must be (2, nt, nh, nw), got {tuple(deformation_grid.shape)}")
    return as_f32(deformation_grid, dev)


def correct_motion(
    image: torch.Tensor,
    deformation_grid: torch.Tensor,
    pixel_spacing: float,
    grad: bool = False,
    grid_type: str = "catmull_rom",
    device: torch.device = None,
    _mean_std: torch.Tensor | None = None,
) -> torch.Tensor:
    """(t, h, w) movie warped by a (yx, nt, nh, nw) Angstrom field -> (t, h, w).

    Reference: correct_motion.py:18-78 (spline -> 10x lattice -> bicubic lattice lookup ->
    bicubic frame gather; quirk Q17).  ``grad`` is accepted for signature compatibility; like
    the reference the result is detached.
    """
    dev = resolve_device(image, device)
    movie = _movie(image, dev)
    field = _field(deformation_grid, dev)
    t = movie.shape[0]
    gh, gw = field.shape[-2:]
    lattice = _ops.spline_lattice(field, grid_kind(grid_type), t, 10 * gh, 10 * gw)
    out = torch.empty_like(movie)
    # _mean_std (internal): warp the normalised movie (image - mean) / std without materialising it
    _ops.warp_lattice(movie, lattice, pixel_spacing, mean_std=_mean_std, out_stack=out)
    return out


def correct_motion_fast(
    image: torch.Tensor,
    deformation_grid: torch.Tensor,
    device: torch.device = None,
    _mean_std: torch.Tensor | None = None,
) -> torch.Tensor:
    """Rigid per-frame Fourier phase shift for a (2, t, 1, 1) field.

    Reference: correct_motion.py:430-498, including quirk Q2: the field is NEGATED IN PLACE (the
    reference's ``shifts`` is a view of the caller's tensor) and its values are used as pixels."""
    from . import _fourier
    from ._lib import call, ptr, stream_ptr

    if tuple(deformation_grid.shape[-2:]) != (1, 1):
        raise ValueError(
            f"Expected single patch deformation field with shape (2, t, 1, 1), "
            f"but got shape {tuple(deformation_grid.shape)}. "
            f"Final two dimensions must be (1, 1) for single patch correction."
        )
    dev = resolve_device(image, device)
    movie = _movie(image, dev)
    t, h, w = movie.shape
    moved = deformation_grid.to(dev)
    moved *= -1  # Q2: the reference negates the caller's tensor in place ...
    if moved.data_ptr() != deformation_grid.data_ptr():
        with torch.no_grad():
            deformation_grid.mul_(-1)  # ... also when .to() had to copy it to the device (CPU callers)
    field = moved.detach().to(torch.float32).contiguous()
    plan = _fourier.BandPlan(h, w, dev, full=True)
    if _fourier.query("tmc_fourier_shift_frames_supported", h, w):
        return plan.shift_frames(movie, _mean_std, field)
    spec = plan.forward(movie, _mean_std, None, 0, h, _fourier.frame_pair_jobs(t, dev), job_mode=2)
    with torch.cuda.device(dev):
        call("tmc_fourier_shift", ptr(spec), t, h, w, ptr(field), 1.0, stream_ptr(dev))
    out = torch.empty_like(movie)
    plan.inverse_full(spec[:t], out)
    return out


def correct_motion_sum(
    image: torch.Tensor,
    deformation_grid: torch.Tensor,
    pixel_spacing: float,
    grid_type: str = "catmull_rom",
    device: torch.device = None,
    out: torch.Tensor | None = None,
    accumulate: bool = False,
    frame_offset: int = 0,
    total_frames: int | None = None,
) -> torch.Tensor:
    """Fused ``correct_motion(...).sum(dim=0)`` that never materialises the warped stack.

    New (additive) entry point: the reference leaves the frame sum to the caller
    (``examples/ttMotion.py:398``).  ``frame_offset``/``total_frames`` let one rank warp a
    contiguous block of frames of a larger movie (frame-split multi-GPU)."""
    dev = resolve_device(image, device)
    movie = _movie(image, dev)
    field = _field(deformation_grid, dev)
    t, h, w = movie.shape
    gh, gw = field.shape[-2:]
    lattice = _ops.spline_lattice(
        field, grid_kind(grid_type), t, 10 * gh, 10 * gw, frame_offset=frame_offset, total_frames=total_frames
    )
    if out is None:
        out = torch.empty((h, w), dtype=torch.float32, device=dev)
        accumulate = False
    _ops.warp_lattice(movie, lattice, pixel_spacing, out_sum=out, accumulate_sum=accumulate)
    return out


def get_pixel_shifts(
    frame: torch.Tensor,
    pixel_spacing: float,
    frame_deformation_grid: torch.Tensor,
    pixel_grid: torch.Tensor = None,
) -> torch.Tensor:
    """(h, w, yx) px shifts from a (yx, gh, gw) Angstrom lattice.  Reference: correct_motion.py:132-185.

    ``pixel_grid`` is accepted for signature compatibility (the kernel derives the integer pixel
    coordinates from thread indices)."""
    dev = resolve_device(frame, None)
    h, w = frame.shape[-2:]
    lattice = as_f32(frame_deformation_grid, dev)
    return _ops.pixel_shifts(lattice, h, w, pixel_spacing)


def correct_motion_slow(
    image: torch.Tensor,
    deformation_grid: torch.Tensor,
    grad: bool = False,
    device: torch.device = None,
) -> torch.Tensor:
    """Exact per-pixel Catmull-Rom evaluation of the field (no lattice, no /pixel_spacing).

    Reference: correct_motion.py:320-427."""
    dev = resolve_device(image, device)
    movie = _movie(image, dev)
    field = _field(deformation_grid, dev)
    t, h, w = movie.shape
    out = torch.empty_like(movie)
    # one frame at a time bounds the (h, w, 3) + (h, w, 2) temporaries
    for f in range(t):
        tyx = _ops.pixel_tyx(h, w, 1, dev, frame_offset=f, total_frames=t)
        shifts = _ops.spline_eval(field, 0, tyx)  # (1, h, w, 2)
        out[f : f + 1] = _ops.warp_dense_shifts(movie[f : f + 1], shifts)
    return out


def _grid_data(grid) -> tuple[torch.Tensor, int]:
    """Accept a spline-grid module (anything with ``.data`` of shape (2, nt, nh, nw)) or a tensor."""
    data = grid.data if hasattr(grid, "data") and not isinstance(grid, torch.Tensor) else grid
    kind = getattr(grid, "grid_kind", None)
    if kind is None:
        name = type(grid).__name__.lower()
        kind = 1 if "bspline" in name else 0
    return data, kind


def correct_motion_two_grids(
    image: torch.Tensor,
    new_deformation_grid,
    base_deformation_grid,
    pixel_spacing: float,
    grad: bool = True,
    device: torch.device = None,
) -> torch.Tensor:
    """Warp with shifts = new(tyx) + base(tyx).  Reference: correct_motion.py:188-317.

    With ``grad=True`` the result carries a graph to ``new_deformation_grid``'s coefficients
    (see ``_autograd.WarpTwoGrids``)."""
    dev = resolve_device(image, device)
    movie = _movie(image, dev)
    new_data, new_kind = _grid_data(new_deformation_grid)
    base_data, base_kind = _grid_data(base_deformation_grid)
    if grad and torch.is_tensor(new_data) and new_data.requires_grad:
        from ._autograd import WarpTwoGrids

        return WarpTwoGrids.apply(new_data, movie, as_f32(base_data, dev), float(pixel_spacing), new_kind, base_kind)
    new_f, base_f = as_f32(new_data, dev), as_f32(base_data, dev)
    t = movie.shape[0]
    gh, gw = new_f.shape[-2:]
    lattice = _ops.spline_lattice(new_f, new_kind, t, 10 * gh, 10 * gw, coeffs2=base_f, kind2=base_kind)
    out = torch.empty_like(movie)
    _ops.warp_lattice(movie, lattice, pixel_spacing, out_stack=out)
    return out
