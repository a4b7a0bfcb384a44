"""Restatement of the third-party numerics the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference (``/root/reference/pyproject.toml:36-46``) depends on five TeamTomo
packages by bare name (no version pin, no lock file).  None is installed in the build
image and none can be fetched (no network), so their *published algorithms* are restated
here.  Parity for this layer is UNPINNED by any artefact of the real packages; it is
frozen by the known-answer tests in ``tests/test_oracle_deps.py``.

=====================================  ==================================================
package (PyPI, unpinned)               symbols restated here
=====================================  ==================================================
torch-cubic-spline-grids               ``CubicBSplineGrid3d``, ``CubicCatmullRomGrid3d``
torch-image-interpolation              ``sample_image_2d``, ``array_to_grid_sample``
torch-fourier-shift                    ``fourier_shift_dft_2d``
torch-fourier-filter                   ``b_envelope``, ``bandpass_filter``,
                                       ``dose_weight_movie``
torch-grid-utils                       ``circle``, ``coordinate_grid``, ``fftfreq_grid``
=====================================  ==================================================

Reference call sites: ``correct_motion.py:6-10,106-109,123-127,170-179,488-494``;
``deformation_field_utils.py:6,31-38``; ``estimate_motion_xc.py:6-7,69-88,262-280``;
``estimate_motion_optimizer.py:9-12,123-129,152-176,495-501``; ``utils.py:6,104-112``.
"""

from __future__ import annotations

import sys
import types
from collections.abc import Sequence

import einops
import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# torch-cubic-spline-grids
# --------------------------------------------------------------------------------------

#: Catmull-Rom characteristic matrix; weights = [1, u, u^2, u^3] @ M
CATMULL_ROM_MATRIX = 0.5 * torch.tensor(
    [[0.0, 2.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [2.0, -5.0, 4.0, -1.0], [-1.0, 3.0, -3.0, 1.0]]
)
#: uniform cubic B-spline characteristic matrix
BSPLINE_MATRIX = (1.0 / 6.0) * torch.tensor(
    [[1.0, 4.0, 1.0, 0.0], [-3.0, 0.0, 3.0, 0.0], [3.0, -6.0, 3.0, 0.0], [-1.0, 3.0, -3.0, 1.0]]
)


def _pad_axis_linear(grid: torch.Tensor, dim: int) -> torch.Tensor:
    """Append one phantom node at each end of ``dim`` by linear extrapolation.

    A singleton axis is first repeated to two identical nodes (constant along it).
    """
    if grid.shape[dim] == 1:
        grid = torch.cat([grid, grid], dim=dim)
    n = grid.shape[dim]
    first = grid.narrow(dim, 0, 1)
    second = grid.narrow(dim, 1, 1)
    last = grid.narrow(dim, n - 1, 1)
    penultimate = grid.narrow(dim, n - 2, 1)
    start = 2 * first - second
    end = 2 * last - penultimate
    return torch.cat([start, grid, end], dim=dim)


def _axis_taps(u: torch.Tensor, n_nodes: int, matrix: torch.Tensor):
    """Per-axis tap indices into the padded axis and the four cubic weights.

    ``u`` (b,) in [0, 1]; ``n_nodes`` is the un-padded node count (>= 2).
    Returns ``idx (b, 4)`` long, ``w (b, 4)``.
    """
    x = u * (n_nodes - 1)
    i = torch.clamp(torch.floor(x), 0, n_nodes - 2)
    tau = x - i
    i = i.long()
    # padded index of original node j is j + 1 -> taps i-1..i+2 are padded i..i+3
    idx = i[:, None] + torch.arange(4, device=u.device)[None, :]
    powers = torch.stack([torch.ones_like(tau), tau, tau * tau, tau * tau * tau], dim=-1)
    w = powers @ matrix.to(powers)
    return idx, w


def evaluate_cubic_grid_3d(data: torch.Tensor, u: torch.Tensor, matrix: torch.Tensor):
    """Evaluate a (c, n0, n1, n2) uniform cubic grid at ``u (..., 3)`` -> ``(..., c)``."""
    lead = u.shape[:-1]
    u = u.reshape(-1, 3).to(data.dtype)
    matrix = matrix.to(u.device)  # the characteristic matrices are module constants (CPU)
    g = data
    for dim in (1, 2, 3):
        g = _pad_axis_linear(g, dim)
    n = [max(s, 2) for s in data.shape[1:]]
    i0, w0 = _axis_taps(u[:, 0], n[0], matrix)
    i1, w1 = _axis_taps(u[:, 1], n[1], matrix)
    i2, w2 = _axis_taps(u[:, 2], n[2], matrix)
    out = []
    chunk = 1 << 18  # the real package also evaluates in mini-batches (no numeric effect)
    for s in range(0, u.shape[0], chunk):
        e = slice(s, s + chunk)
        taps = g[
            :,
            i0[e][:, :, None, None],
            i1[e][:, None, :, None],
            i2[e][:, None, None, :],
        ]  # (c, b, 4, 4, 4)
        val = torch.einsum("cbijk,bi,bj,bk->bc", taps, w0[e], w1[e], w2[e])
        out.append(val)
    out = torch.cat(out, dim=0) if out else data.new_zeros((0, data.shape[0]))
    return out.reshape(*lead, data.shape[0])


class _CubicGrid3d(torch.nn.Module):
    _matrix: torch.Tensor = CATMULL_ROM_MATRIX

    def __init__(self, resolution=(2, 2, 2), n_channels: int = 1):
        super().__init__()
        if isinstance(resolution, int):
            resolution = (resolution,) * 3
        self._data = torch.nn.Parameter(torch.zeros((n_channels, *resolution)))

    @property
    def data(self) -> torch.Tensor:
        return self._data

    @data.setter
    def data(self, value: torch.Tensor) -> None:
        self._data = torch.nn.Parameter(value)

    @property
    def resolution(self):
        return tuple(self._data.shape[1:])

    @property
    def n_channels(self) -> int:
        return self._data.shape[0]

    @classmethod
    def from_grid_data(cls, data: torch.Tensor):
        grid = cls(resolution=tuple(data.shape[-3:]), n_channels=data.shape[0] if data.ndim == 4 else 1)
        if data.ndim == 3:
            data = data[None]
        grid._data = torch.nn.Parameter(data.clone().detach().to(torch.float32))
        return grid

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        return evaluate_cubic_grid_3d(self._data, u, self._matrix)


class CubicCatmullRomGrid3d(_CubicGrid3d):
    _matrix = CATMULL_ROM_MATRIX


class CubicBSplineGrid3d(_CubicGrid3d):
    _matrix = BSPLINE_MATRIX


# --------------------------------------------------------------------------------------
# torch-image-interpolation
# --------------------------------------------------------------------------------------


def array_to_grid_sample(array_coordinates: torch.Tensor, array_shape: Sequence[int]) -> torch.Tensor:
    """Array coordinates (..., d) -> grid_sample coordinates (align_corners=True), flipped."""
    dtype, device = array_coordinates.dtype, array_coordinates.device
    shape = torch.as_tensor(array_shape, dtype=dtype, device=device)
    g = (array_coordinates / (0.5 * shape - 0.5)) - 1
    return torch.flip(g, dims=(-1,))


def sample_image_2d(image: torch.Tensor, coordinates: torch.Tensor, interpolation: str = "bilinear"):
    """Sample (h, w) or (c, h, w) ``image`` at ``coordinates (..., 2)`` (yx, array units).

    ``grid_sample(mode, padding_mode="border", align_corners=True)``; samples whose
    coordinate lies outside ``[0, h-1] x [0, w-1]`` are multiplied by zero.
    """
    if interpolation not in ("nearest", "bilinear", "bicubic"):
        raise ValueError(f"unsupported interpolation {interpolation}")
    had_channels = image.ndim == 3
    if not had_channels:
        image = image[None]
    c, h, w = image.shape
    lead = coordinates.shape[:-1]
    coords = coordinates.reshape(-1, 2).to(image.dtype)
    n = coords.shape[0]
    grid = array_to_grid_sample(coords, (h, w)).reshape(1, 1, n, 2)
    samples = F.grid_sample(
        image[None], grid, mode=interpolation, padding_mode="border", align_corners=True
    )  # (1, c, 1, n)
    samples = samples.reshape(c, n).transpose(0, 1)  # (n, c)
    limit = torch.as_tensor([h - 1, w - 1], dtype=coords.dtype, device=coords.device)
    inside = torch.logical_and(coords >= 0, coords <= limit).all(dim=-1)
    samples = samples * inside[:, None].to(samples.dtype)
    samples = samples.reshape(*lead, c)
    if not had_channels:
        samples = samples[..., 0]
    return samples


# --------------------------------------------------------------------------------------
# torch-grid-utils
# --------------------------------------------------------------------------------------


def coordinate_grid(image_shape, center=None, norm: bool = False, device=None) -> torch.Tensor:
    """(h, w, 2) float32 yx index grid (built with numpy on the host, then moved)."""
    idx = np.indices(tuple(int(s) for s in image_shape)).astype(np.float32)
    grid = torch.as_tensor(idx, device=device)
    grid = einops.rearrange(grid, "d ... -> ... d")
    if center is not None:
        grid = grid - torch.as_tensor(center, dtype=grid.dtype, device=grid.device)
    if norm:
        grid = einops.reduce(grid**2, "... d -> ...", reduction="sum") ** 0.5
    return grid


def fftfreq_grid(image_shape, rfft: bool, fftshift: bool = False, norm: bool = False, device=None):
    """(h, w[, 2]) grid of DFT sample frequencies in cycles/px."""
    h, w = image_shape
    fy = torch.fft.fftfreq(h, device=device)
    fx = torch.fft.rfftfreq(w, device=device) if rfft else torch.fft.fftfreq(w, device=device)
    if fftshift:
        fy = torch.fft.fftshift(fy)
        if not rfft:
            fx = torch.fft.fftshift(fx)
    yy = einops.repeat(fy, "h -> h w", w=len(fx))
    xx = einops.repeat(fx, "w -> h w", h=len(fy))
    grid = einops.rearrange([yy, xx], "f h w -> h w f")
    if norm:
        grid = einops.reduce(grid**2, "h w f -> h w", reduction="sum") ** 0.5
    return grid


def circle(radius: float, image_shape, center=None, smoothing_radius: float = 0, device=None):
    """Soft-edged disc: 1 inside ``radius``; cosine roll-off over ``smoothing_radius`` px
    measured by the Euclidean distance transform of the complement (scipy, on the host)."""
    from scipy import ndimage as ndi

    if isinstance(image_shape, int):
        image_shape = (image_shape, image_shape)
    image_shape = tuple(int(s) for s in image_shape)
    if center is None:
        center = tuple(s // 2 for s in image_shape)
    if device is None:
        device = torch.get_default_device()  # bench.py's ATen-on-CUDA baseline runs this module under torch.device("cuda")
    with torch.device("cpu"):  # the distance transform is scipy's, on the host
        distances = coordinate_grid(image_shape, center=center, norm=True, device=None)
        mask = distances < radius
        if smoothing_radius == 0:
            return mask.float().to(device)
        edt = ndi.distance_transform_edt(torch.logical_not(mask).numpy())
        edt = torch.as_tensor(edt).float()
        idx = torch.logical_and(edt > 0, edt <= smoothing_radius)
        out = mask.float()
        out[idx] = torch.cos((torch.pi / 2) * (edt[idx] / smoothing_radius))
    return out.to(device)


# --------------------------------------------------------------------------------------
# torch-fourier-shift
# --------------------------------------------------------------------------------------


def fourier_shift_dft_2d(dft, image_shape, shifts, rfft: bool, fftshifted: bool):
    """``dft * exp(-2 pi i (f_y s_y + f_x s_x))``: moves image content by +shifts (px)."""
    grid = fftfreq_grid(image_shape, rfft=rfft, fftshift=fftshifted, norm=False, device=dft.device)
    shifts = torch.as_tensor(shifts, dtype=torch.float32, device=dft.device)
    shifts = einops.rearrange(shifts, "... yx -> ... 1 1 yx")
    angles = einops.reduce(-2 * torch.pi * grid * shifts, "... h w yx -> ... h w", reduction="sum")
    return dft * torch.complex(torch.cos(angles), torch.sin(angles))


# --------------------------------------------------------------------------------------
# torch-fourier-filter
# --------------------------------------------------------------------------------------


def b_envelope(B: float, image_shape, pixel_size: float, rfft: bool, fftshift: bool, device=None):
    """``exp(-B (f / pixel_size)^2 / 4)``, f = |(fftfreq_h, [r]fftfreq_w)| cycles/px."""
    f = fftfreq_grid(image_shape, rfft=rfft, fftshift=fftshift, norm=True, device=device)
    f = f / pixel_size
    return torch.exp(-(B * f**2) / 4)


def bandpass_filter(low, high, falloff, image_shape, rfft: bool, fftshift: bool, device=None):
    """1 where ``low < f <= high``; cosine roll-off of width ``falloff`` outside."""
    f = fftfreq_grid(image_shape, rfft=rfft, fftshift=fftshift, norm=True, device=device)
    low = torch.as_tensor(low, dtype=torch.float32, device=f.device)
    high = torch.as_tensor(high, dtype=torch.float32, device=f.device)
    band = torch.logical_and(f > low, f <= high)
    out = band.float()
    if falloff > 0:
        outer = torch.logical_and(f > low - falloff, f <= high + falloff)
        soft = torch.logical_and(outer, ~band)
        d = torch.minimum((f[soft] - low).abs(), (f[soft] - high).abs())
        out[soft] = torch.cos((d / falloff) * (torch.pi / 2))
    return out


def critical_exposure(fft_freq: torch.Tensor, voltage: float = 300.0) -> torch.Tensor:
    """Grant & Grigorieff (2015): N_e(k) = 0.24499 k^-1.6649 + 2.8141 (x0.8 at 200 kV)."""
    scale = 1.0 if voltage >= 300 else 0.8
    eps = 1e-10
    return scale * (0.24499 * torch.clamp(fft_freq, min=eps) ** (-1.6649) + 2.8141)


def dose_weight_movie(
    movie_dft,
    image_shape,
    pixel_size: float,
    pre_exposure: float = 0.0,
    dose_per_frame: float = 1.0,
    voltage: float = 300.0,
    crit_exposure_bfactor: float = -1,
    rfft: bool = True,
    fftshift: bool = False,
):
    """Per-frame exposure filter q_t = exp(-N_t / (2 N_e)), N_t cumulative dose at the END of
    frame t, normalised by sqrt(sum_t q_t^2).  PARITY UNPINNED (examples-only in the reference:
    ``examples/ttMotion.py:331-351``)."""
    f = fftfreq_grid(image_shape, rfft=rfft, fftshift=fftshift, norm=True, device=movie_dft.device)
    k = f / pixel_size
    if crit_exposure_bfactor == -1:
        ne = critical_exposure(k, voltage)
    else:
        ne = 2.0 / (crit_exposure_bfactor * torch.clamp(k, min=1e-10) ** 2)
    n_frames = movie_dft.shape[0]
    dose = pre_exposure + dose_per_frame * torch.arange(1, n_frames + 1, device=movie_dft.device)
    q = torch.exp(-0.5 * dose[:, None, None] / ne[None])
    q = q / torch.sqrt(torch.sum(q**2, dim=0, keepdim=True))
    return movie_dft * q


# --------------------------------------------------------------------------------------
# stand-in installation (lets the UNMODIFIED reference source import these names)
# --------------------------------------------------------------------------------------


def install_stand_ins() -> None:
    """Register this module's restatements under the third-party module names."""

    def _mod(name: str, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    _mod(
        "torch_cubic_spline_grids",
        CubicBSplineGrid3d=CubicBSplineGrid3d,
        CubicCatmullRomGrid3d=CubicCatmullRomGrid3d,
    )
    gsu = _mod("torch_image_interpolation.grid_sample_utils", array_to_grid_sample=array_to_grid_sample)
    _mod("torch_image_interpolation", sample_image_2d=sample_image_2d, grid_sample_utils=gsu)
    _mod("torch_fourier_shift", fourier_shift_dft_2d=fourier_shift_dft_2d)
    env = _mod("torch_fourier_filter.envelopes", b_envelope=b_envelope)
    bp = _mod("torch_fourier_filter.bandpass", bandpass_filter=bandpass_filter)
    dw = _mod("torch_fourier_filter.dose_weight", dose_weight_movie=dose_weight_movie)
    _mod("torch_fourier_filter", envelopes=env, bandpass=bp, dose_weight=dw)
    _mod("torch_grid_utils", circle=circle, coordinate_grid=coordinate_grid, fftfreq_grid=fftfreq_grid)
