"""Host-side planning code of the product package against the oracle (CPU only, no kernels)."""

import random

import numpy as np
import pytest
import torch

from oracle import deps
from oracle import reference_path as rp
from torch_motion_correction_b200 import _fourier
from torch_motion_correction_b200.data_io import read_deformation_field_from_csv, write_deformation_field_to_csv
from torch_motion_correction_b200.distributed import frame_range
from torch_motion_correction_b200.estimate_motion_optimizer import _shuffled_batches
from torch_motion_correction_b200.estimate_motion_xc import _aliasing_schedule
from torch_motion_correction_b200.optimization_state import OptimizationTracker
from torch_motion_correction_b200.patch_grid import patch_centers_1d, patch_grid_centers


@pytest.mark.parametrize(
    "length,patch,expect",
    [  # SURVEY.md Appendix C (computed with the reference's own code)
        (64, 32, [16, 47]),
        (512, 128, [64, 141, 217, 294, 370, 447]),
        (512, 256, [128, 383]),
        (4096, 1024, [512, 1126, 1740, 2355, 2969, 3583]),
    ],
)
def test_patch_centres_match_reference_table(length, patch, expect):
    assert patch_centers_1d(length, patch, patch // 2).tolist() == expect


@pytest.mark.parametrize("shape,patch", [((5, 64, 64), 32), ((10, 512, 512), 128), ((40, 959, 927), 128), ((3, 100, 4096), 64), ((2, 32, 32), 32)])
def test_patch_grid_centres_match_oracle(shape, patch):
    want = rp.patch_grid_centers(shape, (1, patch, patch), (1, patch // 2, patch // 2))
    got = patch_grid_centers(shape, (1, patch, patch), (1, patch // 2, patch // 2))
    assert got.dtype == torch.int64 and torch.equal(got, want)
    assert torch.equal(patch_grid_centers(shape[1:], (patch, patch), (patch // 2, patch // 2)), want[0, :, :, 1:])
    with pytest.raises(ValueError):
        patch_grid_centers(shape, (patch, patch), (patch // 2, patch // 2))


@pytest.mark.parametrize("t", [2, 5, 40, 50, 51, 60, 75, 120])
def test_aliasing_schedule_matches_cache_model(t):
    """Quirk Q1 incl. the eviction regime (t > 50) against the oracle's model of LazyPatchGrid."""
    ref = t // 2
    want = rp.q1_schedule(t, "mean_except_current", ref)
    offsets, deltas = _aliasing_schedule(t, "mean_except_current", ref)
    state = [0] * t
    for k in range(t):
        for d in deltas[offsets[k] : offsets[k + 1]]:
            j = abs(d) - 1
            state[j] = 1 if d > 0 else 0
        expect = list(want[k])
        expect[k] = 0
        assert state == expect, k
    power = _aliasing_schedule(t, "middle_frame", ref)
    want = rp.q1_schedule(t, "middle_frame", ref)
    assert [power[k] for k in range(t) if k != ref] == [want[k] for k in sorted(want)]
    if t <= 50:
        assert [power[k] for k in range(t) if k != ref] == list(range(1, t))


@pytest.mark.parametrize("n,px,fr", [(1024, 1.0, (300, 10)), (1024, 0.83, (300, 10)), (512, 0.83, (300, 10)), (128, 1.0, (300, 10)), (96, 1.3, (120, 6)), (32, 1.3, (120, 6)), (64, 5.0, (300, 10))])
def test_band_box_contains_the_pass_band(n, px, fr):
    low, high = _fourier.band_edges(fr, px)
    l2, h2 = rp.band_edges(fr, px)
    assert low == float(l2) and high == float(h2)
    band = rp.prepare_bandpass_filter(fr, (n, n), px)
    kmax = max(_fourier._max_index_within(n, high), 0)
    ky, kx = band.nonzero(as_tuple=True)
    if len(kx):
        signed = torch.where(ky >= (n + 1) // 2, ky - n, ky)
        assert int(kx.max()) <= kmax and int(signed.abs().max()) <= kmax
        assert int(kx.max()) == min(kmax, n // 2)  # the box is tight along the axes
    # pass-band sizes quoted in SURVEY.md Appendix D.8
    sizes = {(1024, 1.0): 16549, (1024, 0.83): 11402, (512, 0.83): 2877, (128, 1.0): 266}
    if (n, px) in sizes:
        assert int(band.sum()) == sizes[(n, px)]


def test_shuffled_batches_follow_the_reference_iterator_order():
    centers = rp.patch_grid_centers((4, 160, 160), (1, 32, 32), (1, 16, 16))
    g = centers.shape[1] * centers.shape[2]
    random.seed(99)
    want = [sel for sel, _ in rp.patch_batches(centers, (4, 160, 160), 8, True)]
    random.seed(99)
    got = _shuffled_batches(g, 8)
    assert got == want
    assert sorted(i for b in got for i in b) == list(range(g))
    assert [len(b) for b in got][-1] == (g % 8 or 8)


def test_frame_ranges_partition_the_movie():
    for t, world in [(40, 8), (40, 3), (7, 8), (60, 4), (1, 1)]:
        blocks = [frame_range(t, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == t
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_csv_round_trip(tmp_path):
    field = torch.randn((2, 4, 3, 5))
    path = tmp_path / "sub" / "field.csv"
    write_deformation_field_to_csv(field, path)
    header = path.read_text().splitlines()[0]
    assert header == "t,h,w,y_shift,x_shift"
    back = read_deformation_field_from_csv(path)
    assert back.shape == field.shape and torch.equal(back, field)


def test_optimization_tracker(tmp_path):
    tr = OptimizationTracker(sample_every_n_steps=2, total_steps=5)
    assert [tr.sample_this_step(i) for i in range(5)] == [True, False, True, False, True]
    tr.add_checkpoint(torch.zeros((2, 1, 1, 1)), 0.5, 0)
    d = tr.as_dict()
    assert d["optimization_checkpoints"][0]["loss"] == 0.5 and d["total_steps"] == 5
    tr.to_json(str(tmp_path / "t.json"))


def test_no_cpu_fallback():
    import torch_motion_correction_b200 as tmc

    if torch.cuda.is_available():
        pytest.skip("only meaningful without a GPU")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tmc.correct_motion(torch.zeros((2, 8, 8)), torch.zeros((2, 2, 1, 1)), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tmc.estimate_global_motion(torch.zeros((2, 16, 16)), 1.0)


def test_compat_alias_exposes_the_reference_names():
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    for name in [n for n in sys.modules if n == "torch_motion_correction" or n.startswith("torch_motion_correction.")]:
        del sys.modules[name]
    try:
        import torch_motion_correction as mod

        # reference src/torch_motion_correction/__init__.py:32-44
        assert sorted(mod.__all__) == sorted([
            "estimate_local_motion", "correct_motion", "correct_motion_two_grids", "correct_motion_fast", "correct_motion_slow",
            "get_pixel_shifts", "evaluate_deformation_field", "estimate_global_motion",
            "estimate_motion_cross_correlation_patches", "write_deformation_field_to_csv", "read_deformation_field_from_csv",
        ])
        assert all(callable(getattr(mod, n)) for n in mod.__all__)
        from torch_motion_correction.correct_motion import correct_motion, correct_motion_fast  # noqa: F401
        from torch_motion_correction.estimate_motion_optimizer import estimate_local_motion  # noqa: F401
        from torch_motion_correction.estimate_motion_xc import estimate_global_motion  # noqa: F401
    finally:
        sys.path.remove(os.path.join(root, "compat"))
        for name in [n for n in sys.modules if n == "torch_motion_correction" or n.startswith("torch_motion_correction.")]:
            del sys.modules[name]
