"""The whole pipeline (global estimate -> patch cross-correlation -> spline optimiser -> warp-and-sum) on awkward
movie shapes against the oracle: two frames, odd / prime side lengths (Bluestein transforms, the global-memory warp
kernel for rows that are not 16-byte aligned), frames barely larger than one patch, non-square frames and patches that
do not divide the frame.  Tolerances: BASELINE.json north star (shifts <= 0.01 px, sums <= 1e-4 relative L2)."""

import random

import pytest
import torch

import torch_motion_correction_b200 as tmc
from oracle import reference_path as rp

pytestmark = pytest.mark.gpu

SHIFT_PX = 0.01
SUM_REL = 1e-4

CASES = [
    # t, h, w, patch, pixel spacing, frequency range
    (2, 96, 96, 64, 1.0, (60, 4)),        # two frames: leave-one-out reference = the other frame
    (3, 131, 97, 64, 1.2, (60, 4)),       # prime sides: Bluestein whole-frame transforms, unaligned rows
    (5, 70, 150, 64, 0.9, (50, 4)),       # barely taller than a patch, wide
    (4, 257, 258, 96, 1.1, (80, 5)),      # patch side that is not a power of two, W % 4 == 2
    (7, 200, 120, 48, 1.0, (40, 3)),      # many small patches
]


@pytest.mark.parametrize("t,h,w,p,px,fr", CASES)
def test_pipeline_on_awkward_shapes(t, h, w, p, px, fr):
    dev = torch.device("cuda:0")
    movie, _ = rp.synthetic_movie(t, h, w, seed=t * 1000 + h, noise=0.4, drift=2.0, local=0.4, sigma_f=0.12)
    # estimators, stage by stage, each against the oracle on the oracle's own input to the stage
    g_want = rp.estimate_global_motion(movie, px, frequency_range=fr)
    g_got = tmc.estimate_global_motion(movie.to(dev), px, frequency_range=fr)
    assert float((g_got.cpu() - g_want).abs().max()) <= 1e-5
    f_want, pos_want = rp.estimate_motion_cross_correlation_patches(movie, px, patch_sidelength=p, frequency_range=fr)
    f_got, pos_got = tmc.estimate_motion_cross_correlation_patches(movie.to(dev), px, patch_sidelength=p, frequency_range=fr)
    assert torch.equal(pos_got.cpu(), pos_want)
    assert float((f_got.cpu() - f_want).abs().max()) <= SHIFT_PX * px
    res = (min(t, 3), 3, 3)
    random.seed(7)
    l_want = rp.estimate_local_motion(movie, px, (p, p), res, f_want.clone(), n_iterations=4, grid_type="bspline", frequency_range=fr)
    random.seed(7)
    l_got = tmc.estimate_local_motion(movie.to(dev), px, (p, p), res, f_want.clone().to(dev), n_iterations=4, grid_type="bspline",
                                      frequency_range=fr)
    assert float((l_got.cpu() - l_want).abs().max()) <= SHIFT_PX * px
    # correction: fused sum, the stack, and the rigid Fourier-shift route
    want = rp.correct_motion(movie, l_want, px, grid_type="bspline")
    total = tmc.correct_motion_sum(movie.to(dev), l_want.to(dev), px, grid_type="bspline")
    assert float(torch.linalg.norm(total.cpu() - want.sum(dim=0)) / torch.linalg.norm(want.sum(dim=0))) <= SUM_REL
    stack = tmc.correct_motion(movie.to(dev), l_want.to(dev), px, grid_type="bspline")
    assert float(torch.linalg.norm(stack.cpu() - want) / torch.linalg.norm(want)) <= SUM_REL
    fast_want = rp.correct_motion_fast(movie, g_want.clone() / px)
    fast_got = tmc.correct_motion_fast(movie.to(dev), (g_want / px).to(dev))
    assert float(torch.linalg.norm(fast_got.cpu() - fast_want) / torch.linalg.norm(fast_want)) <= SUM_REL


def test_pipeline_driver_on_an_awkward_shape():
    """estimate_motion / motion_correct (the additive drivers) against the oracle's restatement of the same chain."""
    dev = torch.device("cuda:0")
    t, h, w, p, px, fr = 6, 190, 230, 64, 1.1, (60, 4)
    movie, _ = rp.synthetic_movie(t, h, w, seed=99, noise=0.4, drift=3.0, local=0.4, sigma_f=0.12)
    want, _ = rp.estimate_motion_pipeline(movie, px, p, frequency_range=fr)
    got, _ = tmc.estimate_motion(movie.to(dev), px, patch_sidelength=p, frequency_range=fr, n_iterations=0)
    assert float((got.cpu() - want).abs().max()) <= SHIFT_PX * px
    total, field = tmc.motion_correct(movie.to(dev), px, patch_sidelength=p, frequency_range=fr, n_iterations=0)
    ref = rp.correct_motion(movie, want, px, grid_type="bspline").sum(dim=0)
    assert float(torch.linalg.norm(total.cpu() - ref) / torch.linalg.norm(ref)) <= 5e-3  # fields agree to 0.01 px only
    assert torch.equal(field, got)
