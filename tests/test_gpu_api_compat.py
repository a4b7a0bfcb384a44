"""API-compatibility suite: what the reference's own tests assert (shapes, types, devices, errors, the
four zero-field identities, autograd hooks; reference tests/test_estimate_motion.py and
tests/test_correct_motion.py), run through the drop-in alias ``import torch_motion_correction``."""

import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compat"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tmc():
    for name in [n for n in sys.modules if n == "torch_motion_correction" or n.startswith("torch_motion_correction.")]:
        del sys.modules[name]
    import torch_motion_correction as mod

    assert mod.__file__.startswith(os.path.join(ROOT, "compat"))
    return mod


@pytest.fixture(scope="module")
def movie():
    """5x64x64 Gaussian blob drifting +2 px/frame in y, +1 px/frame in x (the reference's fixture)."""
    t, h, w = 5, 64, 64
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    frames = [torch.exp(-((y - (h // 2 + 2 * k)) ** 2 + (x - (w // 2 + k)) ** 2) / 200.0) for k in range(t)]
    return torch.stack(frames)


def test_public_names(tmc):
    for name in ("estimate_local_motion", "correct_motion", "correct_motion_two_grids", "correct_motion_fast",
                 "correct_motion_slow", "get_pixel_shifts", "evaluate_deformation_field", "estimate_global_motion",
                 "estimate_motion_cross_correlation_patches", "write_deformation_field_to_csv",
                 "read_deformation_field_from_csv"):
        assert name in tmc.__all__ and callable(getattr(tmc, name))
    from torch_motion_correction.correct_motion import correct_motion  # noqa: F401
    from torch_motion_correction.estimate_motion_optimizer import estimate_local_motion  # noqa: F401
    from torch_motion_correction.estimate_motion_xc import estimate_global_motion  # noqa: F401


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_estimate_global_motion_shapes(tmc, movie, device):
    f = tmc.estimate_global_motion(image=movie, pixel_spacing=1.0, device=torch.device(device))
    assert isinstance(f, torch.Tensor) and f.shape == (2, 5, 1, 1) and f.device.type == "cuda"
    for kw in (dict(reference_frame=0), dict(b_factor=1000), dict(frequency_range=(200, 20))):
        assert tmc.estimate_global_motion(image=movie, pixel_spacing=1.0, **kw).shape == (2, 5, 1, 1)


def test_estimate_patches_shapes_and_options(tmc, movie):
    f, pos = tmc.estimate_motion_cross_correlation_patches(image=movie, pixel_spacing=1.0, patch_sidelength=32)
    assert f.shape == (2, 5, 2, 2) and pos.shape == (5, 2, 2, 3) and pos.dtype == torch.int64
    for kw in (dict(reference_strategy="middle_frame"), dict(reference_strategy="mean_except_current"),
               dict(sub_pixel_refinement=False), dict(temporal_smoothing=True, smoothing_window_size=3),
               dict(outlier_rejection=True, outlier_threshold=2.0), dict(outlier_rejection=False, temporal_smoothing=False)):
        f, _ = tmc.estimate_motion_cross_correlation_patches(image=movie, pixel_spacing=1.0, patch_sidelength=32, **kw)
        assert f.shape == (2, 5, 2, 2) and torch.isfinite(f).all()
    g = tmc.estimate_global_motion(image=movie, pixel_spacing=1.0)
    f, _ = tmc.estimate_motion_cross_correlation_patches(image=movie, pixel_spacing=1.0, patch_sidelength=32, deformation_field=g)
    assert f.shape == (2, 5, 2, 2)
    with pytest.raises(ValueError, match="Unknown reference_strategy"):
        tmc.estimate_motion_cross_correlation_patches(image=movie, pixel_spacing=1.0, patch_sidelength=32, reference_strategy="nope")


@pytest.mark.parametrize("kw", [
    dict(), dict(optimizer_type="sgd"), dict(grid_type="bspline"), dict(loss_type="ncc"), dict(optimizer_kwargs={"lr": 0.001}),
])
def test_estimate_local_motion_shapes(tmc, movie, kw):
    f = tmc.estimate_local_motion(image=movie, pixel_spacing=1.0, patch_shape=(32, 32), deformation_field_resolution=(2, 2, 2),
                                  initial_deformation_field=None, n_iterations=2, **kw)
    assert f.shape == (2, 2, 2, 2) and torch.isfinite(f).all()
    init = torch.zeros((2, 5, 2, 2))
    f, traj = tmc.estimate_local_motion(movie, 1.0, (32, 32), (2, 2, 2), init, n_iterations=2, return_trajectory=True)
    assert f.shape == (2, 2, 2, 2) and len(traj.checkpoints) == 2


def test_zero_field_identities(tmc, movie):
    """The reference's only numeric assertions."""
    m = movie.cuda()
    zero = torch.zeros((2, 5, 2, 2), device="cuda")
    assert torch.allclose(tmc.correct_motion(m, zero, 1.0), m, atol=0.1)
    assert torch.allclose(tmc.correct_motion_slow(m, zero), m, atol=0.1)
    assert torch.allclose(tmc.correct_motion_fast(m, torch.zeros((2, 5, 1, 1), device="cuda")), m, atol=1e-5)
    from torch_motion_correction.spline_grids import CubicCatmullRomGrid3d

    new = CubicCatmullRomGrid3d(resolution=(5, 2, 2), n_channels=2).to("cuda")
    base = CubicCatmullRomGrid3d(resolution=(5, 2, 2), n_channels=2).to("cuda")
    assert torch.allclose(tmc.correct_motion_two_grids(m, new, base, 1.0, grad=False), m, atol=0.1)


def test_correct_motion_fast_rejects_full_fields(tmc, movie):
    with pytest.raises(ValueError, match="Expected single patch deformation field"):
        tmc.correct_motion_fast(movie, torch.zeros((2, 5, 2, 2)))


def test_estimate_then_correct_is_finite(tmc, movie):
    g = tmc.estimate_global_motion(movie, 1.0)
    out = tmc.correct_motion(movie, g, 1.0)
    assert out.shape == movie.shape and torch.isfinite(out).all() and not out.requires_grad
    f, _ = tmc.estimate_motion_cross_correlation_patches(movie, 1.0, patch_sidelength=32)
    assert torch.isfinite(tmc.correct_motion(movie, f, 1.0, grid_type="bspline")).all()
    shifts = tmc.get_pixel_shifts(movie[0].cuda(), 1.0, torch.zeros((2, 20, 20), device="cuda"), None)
    assert shifts.shape == (64, 64, 2)


def test_two_grids_autograd(tmc, movie):
    from torch_motion_correction.spline_grids import CubicBSplineGrid3d, CubicCatmullRomGrid3d

    for cls in (CubicCatmullRomGrid3d, CubicBSplineGrid3d):
        new = cls(resolution=(3, 2, 2), n_channels=2).to("cuda")
        base = cls.from_grid_data(torch.randn((2, 3, 2, 2)) * 0.5).to("cuda")
        out = tmc.correct_motion_two_grids(movie.cuda(), new, base, 1.0, grad=True)
        assert out.requires_grad and out.shape == movie.shape
        out.sum().backward()
        assert new.data.grad is not None and torch.isfinite(new.data.grad).all() and new.data.grad.abs().sum() > 0
    u = torch.rand((10, 3), device="cuda")
    vals = new(u)
    assert vals.shape == (10, 2)
    vals.sum().backward()
