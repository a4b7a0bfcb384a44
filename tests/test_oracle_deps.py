"""Known-answer tests that freeze the restated third-party layer (``oracle.deps``).

The real packages are not installable here (no network), so these pin the documented
properties the hot path relies on.  CPU only."""

import math

import numpy as np
import pytest
import torch

from oracle import deps


@pytest.mark.parametrize("cls", [deps.CubicBSplineGrid3d, deps.CubicCatmullRomGrid3d])
def test_spline_reproduces_linear_data(cls):
    t, y, x = torch.meshgrid(torch.linspace(0, 1, 4), torch.linspace(0, 1, 5), torch.linspace(0, 1, 3), indexing="ij")
    data = torch.stack([2 * t - y + 3 * x + 1, 0.5 * t + y])
    grid = cls.from_grid_data(data)
    u = torch.rand((100, 3))
    expect = torch.stack([2 * u[:, 0] - u[:, 1] + 3 * u[:, 2] + 1, 0.5 * u[:, 0] + u[:, 1]], dim=-1)
    assert torch.allclose(grid(u), expect, atol=1e-5)


def test_catmull_rom_interpolates_nodes_bspline_approximates():
    data = torch.randn((2, 4, 5, 6))
    nodes = torch.stack(
        torch.meshgrid(torch.linspace(0, 1, 4), torch.linspace(0, 1, 5), torch.linspace(0, 1, 6), indexing="ij"), dim=-1
    )
    cr = deps.CubicCatmullRomGrid3d.from_grid_data(data)(nodes)
    assert torch.allclose(cr.permute(3, 0, 1, 2), data, atol=1e-5)
    bs = deps.CubicBSplineGrid3d.from_grid_data(data)(nodes).permute(3, 0, 1, 2)
    # interior node along the last axis only: (p- + 4 p0 + p+)/6 blend in each axis
    k = torch.tensor([1.0, 4.0, 1.0]) / 6
    blend = torch.einsum("ctyx,t,y,x->c", data[:, 0:3, 1:4, 2:5], k, k, k)
    assert torch.allclose(bs[:, 1, 2, 3], blend, atol=1e-5)


def test_spline_partition_of_unity_and_singleton_axes():
    const = torch.full((2, 3, 1, 1), 2.5)
    for cls in (deps.CubicBSplineGrid3d, deps.CubicCatmullRomGrid3d):
        out = cls.from_grid_data(const)(torch.rand((50, 3)))
        assert torch.allclose(out, torch.full_like(out, 2.5), atol=1e-6)
    # (2, t, 1, 1): constant in space, spline in time
    data = torch.randn((2, 5, 1, 1))
    u = torch.rand((20, 3))
    a = deps.CubicCatmullRomGrid3d.from_grid_data(data)(u)
    u2 = u.clone()
    u2[:, 1:] = torch.rand((20, 2))
    assert torch.allclose(a, deps.CubicCatmullRomGrid3d.from_grid_data(data)(u2), atol=1e-6)


def test_spline_gradient_is_transpose():
    grid = deps.CubicBSplineGrid3d(resolution=(3, 4, 4), n_channels=2)
    u = torch.rand((7, 3))
    out = grid(u)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    # linear map: f(data) = B data  =>  grad = B^T w ; check with a second random data
    d2 = torch.randn_like(grid.data)
    lhs = (deps.evaluate_cubic_grid_3d(d2, u, deps.BSPLINE_MATRIX) * w).sum()
    rhs = (grid.data.grad * d2).sum()
    assert torch.allclose(lhs, rhs, atol=1e-5)


def test_fourier_shift_integer_equals_roll():
    img = torch.randn((3, 16, 20))
    shifts = torch.tensor([[1.0, -2.0], [0.0, 3.0], [-4.0, 5.0]])
    out = torch.fft.irfftn(
        deps.fourier_shift_dft_2d(torch.fft.rfftn(img, dim=(-2, -1)), (16, 20), shifts, rfft=True, fftshifted=False), s=(16, 20)
    )
    for k in range(3):
        assert torch.allclose(out[k], torch.roll(img[k], (int(shifts[k, 0]), int(shifts[k, 1])), dims=(0, 1)), atol=1e-5)


def test_bicubic_sampling_identity_and_outside_zero():
    img = torch.randn((24, 31))
    grid = deps.coordinate_grid((24, 31))
    assert torch.allclose(deps.sample_image_2d(img, grid, "bicubic"), img, atol=1e-5)
    outside = torch.tensor([[-0.01, 3.0], [23.01, 3.0], [5.0, 30.5], [23.0, 30.0]])
    vals = deps.sample_image_2d(img, outside, "bicubic")
    assert torch.equal(vals[:3], torch.zeros(3))
    assert abs(float(vals[3] - img[23, 30])) < 1e-5


def test_filters():
    env = deps.b_envelope(500, (32, 32), 1.5, rfft=True, fftshift=False)
    assert env.shape == (32, 17) and float(env[0, 0]) == 1.0
    f = 0.25 / 1.5
    assert abs(float(env[0, 8]) - math.exp(-500 * f * f / 4)) < 1e-6
    band = deps.bandpass_filter(0.1, 0.25, 0, (32, 32), rfft=True, fftshift=False)
    assert float(band[0, 0]) == 0.0  # DC removed
    assert float(band[0, 8]) == 1.0  # f == high is inside (<=)
    assert float(band[0, 9]) == 0.0
    fx = torch.fft.rfftfreq(32)
    assert float(band[0, int((fx > 0.1).nonzero()[0])]) == 1.0  # f == low is outside (>)


def test_circle_mask():
    m = deps.circle(8, (32, 32), smoothing_radius=4)
    assert m.shape == (32, 32) and float(m[16, 16]) == 1.0 and float(m[0, 0]) == 0.0
    assert float(m[16, 16 + 7]) == 1.0  # strictly inside radius
    # first pixel outside the disc: EDT distance 1 -> cos(pi/2 * 1/4)
    assert abs(float(m[16, 16 + 8]) - math.cos(math.pi / 8)) < 1e-6
    assert float(m[16, 16 + 12]) == 0.0 or abs(float(m[16, 16 + 11]) - math.cos(math.pi / 2)) < 1e-6
    hard = deps.circle(8, (32, 32), smoothing_radius=0)
    assert set(np.unique(hard.numpy())) == {0.0, 1.0}


def test_dose_weight_known_answers():
    """Exposure filter restated from Grant & Grigorieff: q = 1 at DC, monotonically stronger damping of late frames at
    high resolution, unit total power (sum_t q_t^2 / norm^2 == 1), and the example-script pipeline is linear."""
    from oracle import reference_path as rp

    t, h, w = 5, 32, 48
    ones = torch.ones((t, h, w // 2 + 1), dtype=torch.complex64)
    q = deps.dose_weight_movie(ones, (h, w), 1.0, 0.0, 2.0, 300.0, -1, True, False).real
    assert torch.allclose((q**2).sum(dim=0), torch.ones((h, w // 2 + 1)), atol=1e-5)
    assert torch.allclose(q[:, 0, 0], torch.full((t,), 1 / t**0.5), atol=1e-6)  # DC: every frame weighted equally
    assert bool((q[0, 5:, 5:] > q[-1, 5:, 5:]).all())  # early frames carry the high resolution
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn((t, h, w), generator=g), torch.randn((t, h, w), generator=g)
    lhs = rp.dose_weight(a + 2 * b, 1.1, 0.5, 1.0)
    assert torch.allclose(lhs, rp.dose_weight(a, 1.1, 0.5, 1.0) + 2 * rp.dose_weight(b, 1.1, 0.5, 1.0), atol=1e-4)


# ---- movie preparation (examples/ttMotion.py:90-202 restated in oracle.reference_path.prepare_movie) ----


def test_prepare_movie_restatement_follows_the_example():
    import numpy as np

    from oracle import reference_path as rp

    rng = np.random.default_rng(7)
    movie = rng.poisson(20.0, size=(3, 24, 31)).astype(np.uint16)
    gain = (1.0 + 0.1 * rng.standard_normal((24, 31))).astype(np.float32)
    # gain_correct (:123) and set_frames_mean_zero (:195-198), transliterated
    want = movie.astype(np.float32) * gain
    want = want - np.mean(want, axis=(1, 2), keepdims=True)
    got, n_hot = rp.prepare_movie(movie, gain=gain, zero_frame_means=True)
    assert n_hot == 0 and got.dtype == np.float32
    assert np.allclose(got, want, atol=1e-5)
    # remove_hot_pixels (:141-176): the mask is the reference's, the replacement one of the neighbours
    movie[1, 10, 12] = 4000
    movie[2, 0, 0] = 3000
    f32 = movie.astype(np.float32)
    got, n_hot = rp.prepare_movie(movie, hot_pixel_threshold=10.0)
    hot = np.zeros_like(f32, dtype=bool)
    for f in range(3):
        m, s = f32[f].mean(), f32[f].std()
        hot[f] = (f32[f] > m + 10.0 * s) | (f32[f] < m - 10.0 * s)
    assert n_hot == int(hot.sum()) == 2
    assert np.array_equal(got[~hot], f32[~hot])
    assert got[1, 10, 12] in f32[1, 9:12, 11:14] and got[1, 10, 12] != 4000
    assert got[2, 0, 0] in (f32[2, 0, 1], f32[2, 1, 0], f32[2, 1, 1])
    # reproducible
    again, _ = rp.prepare_movie(movie, hot_pixel_threshold=10.0)
    assert np.array_equal(got, again)


def test_estimate_motion_pipeline_restatement_composes_before_smoothing():
    """global + patch residuals, smoothed after the composition, then the example's refinement loop: shapes, the joint
    mean (quirk Q4) and the early stop."""
    from oracle import reference_path as rp

    movie, _ = rp.synthetic_movie(6, 96, 96, seed=4, noise=0.4, drift=2.0, local=0.5, sigma_f=0.07)
    field, hist = rp.estimate_motion_pipeline(movie, 1.2, 32, frequency_range=(100, 5), n_refinements=2, refinement_tolerance=1e9)
    assert field.shape[0] == 2 and field.shape[1] == 6 and len(hist) == 1
    assert abs(float(field.mean())) < 1e-5
    field0, hist0 = rp.estimate_motion_pipeline(movie, 1.2, 32, frequency_range=(100, 5))
    assert hist0 == [] and abs(float(field0.mean())) < 1e-5
