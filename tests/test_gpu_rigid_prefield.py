"""Patch cross-correlation on top of a rigid (2, t, 1, 1) field: whole-pixel fields are served by reading the patch
windows at shifted origins (an integer Fourier shift is a circular roll) instead of a Fourier-shift pass over the stack
-- same result as the reference's pre-correction route (estimate_motion_xc.py:232-257); the composition of global and
patch fields in the pipeline; even Savitzky-Golay windows (frame counts below the window)."""

import numpy as np
import pytest
import torch

import torch_motion_correction_b200 as tmc
from oracle import reference_path as rp
from torch_motion_correction_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _xc(movie, px, p, field, whole, **kw):
    before = dict(_lib.CALLS)
    f, _ = tmc.estimate_motion_cross_correlation_patches(
        movie, px, patch_sidelength=p, deformation_field=field, _whole_pixel_field=whole, **kw)
    shift_calls = sum(_lib.CALLS.get(k, 0) - before.get(k, 0) for k in ("tmc_fourier_shift_frames", "tmc_fourier_shift"))
    return f, shift_calls


# (frames, size, patch, largest shift): soft path (window leaves the frame only where the mask is zero) and wrapped
# reads (shift beyond the mask margin p/8) for the generic (Bluestein / small), power-of-two and polyphase row kernels
CASES = [
    (6, 96, 32, 3), (6, 96, 32, 9), (5, 100, 36, 7),
    (6, 512, 128, 6), (6, 512, 128, 40), (4, 768, 256, 25), (4, 768, 256, 70),
    (4, 2048, 1024, 17), (4, 2048, 1024, 300),
]


@pytest.mark.parametrize("t,n,p,amp", CASES)
@pytest.mark.parametrize("strategy", ["mean_except_current", "middle_frame"])
def test_shifted_windows_equal_the_fourier_shift_route(dev, t, n, p, amp, strategy):
    px = 1.1
    movie, _ = rp.synthetic_movie(t, n, n, seed=n + amp, noise=0.6, drift=min(amp, 4.0), integer_shifts=True, sigma_f=0.08)
    movie = movie.to(dev)
    g = torch.Generator().manual_seed(amp)
    field = torch.randint(-amp, amp + 1, (2, t, 1, 1), generator=g).float()
    field[:, 0] = torch.tensor([amp, -amp]).reshape(2, 1, 1)  # the extremes are always present
    fr = (300.0, 10.0) if p >= 128 else (120.0, 6.0)
    kw = dict(reference_strategy=strategy, frequency_range=fr, temporal_smoothing=False, outlier_rejection=False)
    handed = field.clone().to(dev)
    rolled, calls = _xc(movie, px, p, handed, None, **kw)
    assert calls == 0  # whole numbers are detected: no pass over the stack
    assert torch.equal(handed.cpu(), -field)  # quirk Q2: the caller's tensor is negated in place either way
    handed = field.clone().to(dev)
    shifted, calls = _xc(movie, px, p, handed, False, **kw)
    assert calls >= 1
    assert torch.equal(handed.cpu(), -field)
    err = float((rolled - shifted).abs().max()) / px
    assert err <= 1e-4, err


def test_fractional_field_takes_the_fourier_shift_route(dev):
    px = 1.0
    movie, _ = rp.synthetic_movie(6, 256, 256, seed=5, noise=0.5, drift=3.0, integer_shifts=True, sigma_f=0.08)
    movie = movie.to(dev)
    field = torch.tensor([[1.0, -2.0, 0.5, 3.0, 0.0, 1.0], [0.0, 1.0, 2.0, -1.25, 0.0, 2.0]]).reshape(2, 6, 1, 1).to(dev)
    f, calls = _xc(movie, px, 64, field.clone(), None, frequency_range=(120.0, 6.0))
    assert calls >= 1
    want, _ = rp.estimate_motion_cross_correlation_patches(
        movie.cpu(), px, patch_sidelength=64, frequency_range=(120.0, 6.0), deformation_field=field.clone().cpu())
    assert float((f.cpu() - want).abs().max()) <= 0.01 * px


def test_whole_pixel_route_matches_the_oracle(dev):
    """The oracle restates the reference's route (Fourier shift of every frame, then patch extraction)."""
    px = 1.3
    movie, _ = rp.synthetic_movie(6, 128, 128, seed=17, noise=0.5, drift=3.0, integer_shifts=True, sigma_f=0.07)
    field = torch.tensor([[2.0, -1.0, 0.0, 3.0, -6.0, 1.0], [-3.0, 0.0, 0.0, 5.0, 2.0, -7.0]]).reshape(2, 6, 1, 1)
    fr = (120.0, 6.0)
    want, _ = rp.estimate_motion_cross_correlation_patches(
        movie, px, patch_sidelength=32, frequency_range=fr, deformation_field=field.clone())
    got, calls = _xc(movie.to(dev), px, 32, field.clone().to(dev), None, frequency_range=fr)
    assert calls == 0
    assert float((got.cpu() - want).abs().max()) <= 0.01 * px


@pytest.mark.parametrize("t", [4, 3, 6])
def test_smoothing_window_capped_at_an_even_frame_count(dev, t):
    """Default window 5 on t = 4 frames becomes 4: scipy >= 1.11 (and so the reference) accepts the even window."""
    px = 1.0
    movie, _ = rp.synthetic_movie(t, 128, 128, seed=t, noise=0.5, drift=2.0, sigma_f=0.07)
    fr = (120.0, 6.0)
    want, _ = rp.estimate_motion_cross_correlation_patches(movie, px, patch_sidelength=32, frequency_range=fr)
    got, _ = tmc.estimate_motion_cross_correlation_patches(movie.to(dev), px, patch_sidelength=32, frequency_range=fr)
    assert float((got.cpu() - want).abs().max()) <= 0.01 * px


def test_pipeline_composes_global_and_patch_fields_before_smoothing(dev):
    """estimate_motion = smooth(global + patch residuals) - mean, with the Savitzky-Golay filter applied AFTER the bases
    are swapped.  (After the whole-pixel pre-correction the residuals lie within half a pixel, where the reference's
    sub-pixel refinement is inactive -- the peak sits on the border of the un-shifted correlation image, quirk Q6 -- so
    the cross-correlation stage tracks a smooth drift as a smoothed staircase; the spline optimiser refines it.)"""
    from scipy.signal import savgol_filter

    t, n, px = 24, 1024, 0.83
    g = torch.Generator().manual_seed(1)
    pad = 64
    white = torch.randn((n + 2 * pad, n + 2 * pad), generator=g)
    fy = torch.fft.fftfreq(n + 2 * pad)[:, None]
    fx = torch.fft.rfftfreq(n + 2 * pad)[None, :]
    spec = torch.fft.rfftn(white) * torch.exp(-(fy**2 + fx**2) / (2 * 0.08**2))
    k = torch.arange(t, dtype=torch.float32) - t // 2
    true = torch.stack([0.11 * k, -0.07 * k + 0.002 * k * k], dim=1)  # px, content displacement of frame k
    frames = []
    for i in range(t):
        phase = torch.exp(-2j * np.pi * (fy * float(true[i, 0]) + fx * float(true[i, 1])))
        frame = torch.fft.irfftn(spec * phase, s=white.shape)[pad:-pad, pad:-pad]
        frames.append(frame / frame.std() + 0.3 * torch.randn((n, n), generator=g))
    movie = torch.stack(frames).float().to(dev)
    field, _ = tmc.estimate_motion(movie, px, patch_sidelength=512)

    # the same thing from its parts, smoothed on the host with scipy
    glob = tmc.estimate_global_motion(movie, px)
    assert float((glob[:, :, 0, 0].T.cpu() / px - torch.round(true)).abs().max()) <= 1.0
    pre = (glob / px).clone()
    raw, _ = tmc.estimate_motion_cross_correlation_patches(movie, px, patch_sidelength=512, deformation_field=pre,
                                                           temporal_smoothing=False)
    composed = (raw + glob / px + glob).cpu().numpy()  # raw was accumulated on -(glob / px): swap it for glob
    want = savgol_filter(composed, 5, 1, axis=1)
    want = want - want.mean()
    assert float(np.abs(field.cpu().numpy() - want).max()) <= 2e-4 * px
    # the smoothed staircase stays within half a pixel of the true drift (tracks compared up to a constant per axis)
    truth = true.T * px
    truth = truth - truth.mean(dim=1, keepdim=True)
    got = field.mean(dim=(2, 3)).cpu()
    got = got - got.mean(dim=1, keepdim=True)
    assert float((got - truth).abs().max()) / px <= 0.5


def test_iterative_refinement_matches_the_oracle(dev):
    """estimate_motion(n_refinements=...) -- the example's loop (examples/ttMotion.py:287-329) -- pass by pass against the
    oracle's restatement, early stop included."""
    px, fr = 1.3, (120.0, 6.0)
    movie, _ = rp.synthetic_movie(8, 160, 160, seed=23, noise=0.4, drift=3.0, local=1.0, sigma_f=0.07)
    want, want_hist = rp.estimate_motion_pipeline(movie, px, 32, frequency_range=fr, n_refinements=3, refinement_tolerance=1e-3)
    got, _, hist = tmc.estimate_motion(movie.to(dev), px, patch_sidelength=32, frequency_range=fr, n_refinements=3,
                                       refinement_tolerance=1e-3, return_history=True)
    assert len(hist) == len(want_hist) >= 1
    assert float((got.cpu() - want).abs().max()) <= 0.01 * px
    assert np.allclose(hist, want_hist, atol=2e-3 * px)
    # a loose tolerance stops after the first pass
    _, _, hist1 = tmc.estimate_motion(movie.to(dev), px, patch_sidelength=32, frequency_range=fr, n_refinements=3,
                                      refinement_tolerance=1e3, return_history=True)
    assert len(hist1) == 1


@pytest.mark.parametrize("values", [[2.0, -1.0, 0.0, 3.0, -6.0, 1.0], [2.5, -1.0, 0.0, 3.0, -6.0, 1.0]])
def test_quirk_q2_reaches_a_cpu_callers_tensor(dev, values):
    """correct_motion_fast negates the caller's field in place (quirk Q2): also when the caller's tensor lives on the host
    and had to be copied to the device -- whole-pixel (shifted windows) and fractional (Fourier shift) routes."""
    movie, _ = rp.synthetic_movie(6, 128, 128, seed=17, noise=0.5, drift=3.0, integer_shifts=True, sigma_f=0.07)
    field = torch.tensor([values, values[::-1]]).reshape(2, 6, 1, 1)
    keep = field.clone()
    tmc.estimate_motion_cross_correlation_patches(movie.to(dev), 1.3, patch_sidelength=32, frequency_range=(120.0, 6.0),
                                                  deformation_field=field)
    assert torch.equal(field, -keep)
    field = keep.clone()
    tmc.correct_motion_fast(movie.to(dev), field)
    assert torch.equal(field, -keep)
