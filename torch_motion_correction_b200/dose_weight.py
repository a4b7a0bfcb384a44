"""Dose weighting of an aligned movie (additive API; the reference only does this in its example script,
``examples/ttMotion.py:331-351``, through ``torch_fourier_filter.dose_weight.dose_weight_movie``)."""

from __future__ import annotations

import torch

from . import _fourier
from ._common import as_f32, resolve_device
from ._lib import call, ptr, stream_ptr

#: frames transformed per block: bounds the (frames, ny, nx/2+1) complex64 spectra held at once
_BLOCK_BYTES = 4 << 30


def dose_weight(
    movie: torch.Tensor,
    pixel_size: float,
    pre_exposure: float = 0.0,
    dose_per_frame: float = 1.0,
    voltage: float = 300.0,
    device: torch.device = None,
) -> torch.Tensor:
    """``sum_t irfft2(rfft2(frame_t) * q_t / sqrt(sum_t q_t^2))`` with the Grant & Grigorieff exposure filter
    ``q_t = exp(-N_t / (2 N_e(k)))``, ``N_t = pre_exposure + (t + 1) * dose_per_frame``: the dose-weighted sum
    ``(h, w)`` of an already motion-corrected ``(t, h, w)`` stack.  The filter is applied to the Fourier-space SUM
    (linear), so only one inverse transform is needed."""
    dev = resolve_device(movie, device)
    stack = as_f32(movie, dev)
    t, h, w = stack.shape
    plan = _fourier.BandPlan(h, w, dev, full=True)
    kxn = w // 2 + 1
    acc = torch.zeros((h, kxn, 2), dtype=torch.float32, device=dev)
    den2 = torch.zeros((h, kxn), dtype=torch.float32, device=dev)
    block = max(2, min(t, int(_BLOCK_BYTES // (h * kxn * 8)) // 2 * 2))
    for f0 in range(0, t, block):
        n = min(block, t - f0)
        spec = plan.forward(stack[f0 : f0 + n], None, None, 0, h, _fourier.frame_pair_jobs(n, dev), job_mode=2)[:n]
        with torch.cuda.device(dev):
            call("tmc_dose_weighted_sum", ptr(spec), n, h, w, float(pixel_size), float(pre_exposure), float(dose_per_frame),
                 float(voltage), f0, ptr(acc), ptr(den2), int(f0 + n >= t), stream_ptr(dev))
        del spec
    out = torch.empty((1, h, w), dtype=torch.float32, device=dev)
    plan.inverse_full(acc.reshape(1, h, kxn, 2), out)
    return out[0]
