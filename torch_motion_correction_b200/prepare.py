"""Movie preparation ahead of the estimators (additive API): detector-native pixel types -> fp32, gain multiply,
hot-pixel replacement, per-frame mean removal -- the NumPy pre-processing of the reference's example workflow
(``examples/ttMotion.py:90-202``: ``gain_correct``, ``remove_hot_pixels``, ``set_frames_mean_zero``, and the cast to
float32 at ``:357``) as CUDA kernels behind the C ABI (``csrc/prepare.cu``)."""

from __future__ import annotations

import torch

from ._common import resolve_device
from ._lib import call, ptr, query, stream_ptr

DTYPE_CODES = {torch.uint8: 0, torch.uint16: 1, torch.int16: 2, torch.float16: 3, torch.float32: 4, torch.int8: 5}


def oriented_gain(gain: torch.Tensor, flip_gain: int = 0, rot_gain: int = 0) -> torch.Tensor:
    """The gain map as ``gain_correct`` orients it (examples/ttMotion.py:115-123): flip (1 = flipY, 2 = flipX), then
    ``np.rot90(k=-rot_gain)``."""
    if flip_gain == 1:
        gain = torch.flip(gain, dims=(0,))
    elif flip_gain == 2:
        gain = torch.flip(gain, dims=(1,))
    if rot_gain != 0:
        gain = torch.rot90(gain, k=-int(rot_gain), dims=(0, 1))
    return gain.contiguous()


def prepare_movie(movie: torch.Tensor, gain: torch.Tensor | None = None, hot_pixel_threshold: float | None = None,
                  zero_frame_means: bool = False, device=None, out: torch.Tensor | None = None,
                  max_hot_pixels: int = 1 << 20, return_hot_pixel_count: bool = False):
    """(t, h, w) movie of dtype uint8 / int8 / uint16 / int16 / float16 / float32 -> fp32 device tensor.

    ``gain`` (h, w): multiplied in (``gain_correct``); ``hot_pixel_threshold``: pixels further than that many standard
    deviations from their frame's mean are replaced by a neighbour (``remove_hot_pixels``; the reference picks the
    neighbour with ``np.random.choice``, here a hash of the pixel position picks it, reproducibly); ``zero_frame_means``:
    every frame minus its own mean (``set_frames_mean_zero``).  Same order as the example's ``main``.  One pass over the
    stack converts, applies the gain and accumulates the per-frame moments the other two steps need."""
    if movie.ndim != 3:
        raise ValueError(f"movie must be (t, h, w), got {tuple(movie.shape)}")
    if movie.dtype not in DTYPE_CODES:
        raise TypeError(f"unsupported movie dtype {movie.dtype}: expected one of {sorted(str(d) for d in DTYPE_CODES)}")
    dev = resolve_device(movie, device)
    src = movie.detach().to(dev).contiguous()
    t, h, w = src.shape
    n = h * w
    if out is None:
        out = torch.empty((t, h, w), dtype=torch.float32, device=dev)
    elif out.shape != src.shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 tensor of the movie's shape")
    gain_dev = None
    if gain is not None:
        if tuple(gain.shape) != (h, w):
            raise ValueError(f"gain must be ({h}, {w}), got {tuple(gain.shape)}")
        gain_dev = gain.detach().to(device=dev, dtype=torch.float32).contiguous()
    need_moments = hot_pixel_threshold is not None or zero_frame_means
    moments = torch.empty((t, 2), dtype=torch.float64, device=dev) if need_moments else None
    count = None
    with torch.cuda.device(dev):
        stream = stream_ptr(dev)
        call("tmc_convert_stack", ptr(src), DTYPE_CODES[movie.dtype], t, n, ptr(gain_dev), ptr(out), ptr(moments), stream)
        if hot_pixel_threshold is not None:
            ws = torch.empty((query("tmc_hot_pixel_workspace_bytes", int(max_hot_pixels)),), dtype=torch.uint8, device=dev)
            count = torch.zeros((1,), dtype=torch.int32, device=dev)
            call("tmc_remove_hot_pixels", ptr(out), t, h, w, ptr(moments), float(hot_pixel_threshold), int(max_hot_pixels), ptr(ws),
                 ptr(count), stream)
        if zero_frame_means:
            call("tmc_subtract_frame_means", ptr(out), t, n, ptr(moments), stream)
    if return_hot_pixel_count:
        return out, count
    return out
