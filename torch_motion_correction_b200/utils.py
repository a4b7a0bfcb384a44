"""Normalisation and unit helpers (mirror of the reference's ``utils.py``)."""

from __future__ import annotations

import torch

from . import _fourier, _ops
from ._common import as_f32, resolve_device


def normalize_image(image: torch.Tensor, frac_low: float = 0.25, frac_high: float = 0.75) -> torch.Tensor:
    """(image - mean) / std with scalar statistics of the central box of the whole stack.

    Reference: utils.py:49-84.  The estimators never call this: they fuse the affine into their
    loads; it exists for API compatibility."""
    dev = resolve_device(image, None)
    movie = as_f32(image, dev)
    stats = _ops.stack_stats(movie, frac_low, frac_high)
    return (movie - stats[0]) / stats[1]


def spatial_frequency_to_fftfreq(frequencies, spacing: float) -> torch.Tensor:
    """cycles / unit distance -> cycles / px (utils.py:41-46)."""
    return torch.as_tensor(frequencies, dtype=torch.float32) * spacing


def fftfreq_to_spatial_frequency(frequencies, spacing: float) -> torch.Tensor:
    """cycles / px -> cycles / unit distance (utils.py:33-38)."""
    return torch.as_tensor(frequencies, dtype=torch.float32) * (1 / spacing)


def prepare_bandpass_filter(frequency_range, patch_shape, pixel_spacing: float, refinement_fraction: float = 1.0, device=None):
    """Dense (ph, pw//2+1) hard band-pass table (utils.py:87-114); the kernels use the band-limited
    form built by ``tmc_band_weights`` instead."""
    from ._lib import call, ptr, stream_ptr

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    ph, pw = patch_shape
    low, high = _fourier.band_edges(frequency_range, pixel_spacing)
    out = torch.empty((ph, pw // 2 + 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("tmc_band_weights", ph, pw, ph, pw // 2 + 1, 0, low, high, 1, 0.0, float(pixel_spacing), 0, ptr(out), stream_ptr(dev))
    return out
