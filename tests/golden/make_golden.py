"""Generate the committed golden vectors from the UNMODIFIED reference source.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

It imports ``/root/reference/src/torch_motion_correction`` verbatim on top of the restated
third-party layer (``oracle.deps``; the real packages are not installable here), runs the
reference's own public functions on small seeded movies on the CPU and stores inputs and
outputs in ``tests/golden/*.npz``.  The ``-m "not gpu"`` tests check ``oracle.reference_path``
against these files; the ``-m gpu`` tests check the CUDA path against them as well.
"""

from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import deps, verbatim  # noqa: E402
from oracle.reference_path import synthetic_movie  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def np32(x):
    return x.detach().cpu().numpy()


def smooth_field(seed, shape, scale):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((2, *shape), generator=g) * scale).float()


def case_small(ref):
    """6x96x96 movie: every public entry point once."""
    out = {}
    px = 1.3
    movie, walk = synthetic_movie(6, 96, 96, seed=11, noise=0.5, drift=3.0, local=0.8, sigma_f=0.07)
    out["movie"] = np32(movie)
    out["pixel_spacing"] = np.float64(px)
    fr = (120.0, 6.0)  # wide enough to keep a useful band on 32/48 px patches
    out["frequency_range"] = np.asarray(fr)
    with verbatim.quiet():
        out["global_field"] = np32(ref.estimate_global_motion(movie.clone(), px, frequency_range=fr))
        out["global_field_ref0_b1000"] = np32(
            ref.estimate_global_motion(movie.clone(), px, reference_frame=0, b_factor=1000, frequency_range=fr)
        )
        for strat in ("mean_except_current", "middle_frame"):
            f, pos = ref.estimate_motion_cross_correlation_patches(
                movie.clone(), px, reference_strategy=strat, patch_sidelength=32, frequency_range=fr,
                temporal_smoothing=False, outlier_rejection=False,
            )
            out[f"xc_raw_{strat}"] = np32(f)
            f, pos = ref.estimate_motion_cross_correlation_patches(
                movie.clone(), px, reference_strategy=strat, patch_sidelength=32, frequency_range=fr,
                smoothing_window_size=3, outlier_threshold=1.5,
            )
            out[f"xc_full_{strat}"] = np32(f)
            out["xc_positions"] = np32(pos)
        f, _ = ref.estimate_motion_cross_correlation_patches(
            movie.clone(), px, patch_sidelength=32, frequency_range=fr, sub_pixel_refinement=False,
            temporal_smoothing=False, outlier_rejection=False,
        )
        out["xc_integer"] = np32(f)
        # cumulative paths: (2,t,1,1) field -> correct_motion_fast (Q2), full field -> correct_motion(bspline)
        g0 = torch.as_tensor(out["global_field"]).clone()
        f, _ = ref.estimate_motion_cross_correlation_patches(
            movie.clone(), px, patch_sidelength=32, frequency_range=fr, deformation_field=g0
        )
        out["xc_cumulative_global"] = np32(f)
        out["xc_cumulative_global_field_after"] = np32(g0)  # Q2: negated in place
        f0 = smooth_field(5, (6, 3, 3), 1.5)
        f, _ = ref.estimate_motion_cross_correlation_patches(
            movie.clone(), px, patch_sidelength=32, frequency_range=fr, deformation_field=f0.clone()
        )
        out["xc_cumulative_full_in"] = np32(f0)
        out["xc_cumulative_full"] = np32(f)

        # correction
        field = smooth_field(3, (3, 4, 4), 2.0)
        out["field_344"] = np32(field)
        out["correct_catmull"] = np32(ref.correct_motion(movie, field.clone(), px))
        out["correct_bspline"] = np32(ref.correct_motion(movie, field.clone(), px, grid_type="bspline"))
        field1 = smooth_field(4, (6, 1, 1), 2.5)
        out["field_611"] = np32(field1)
        out["correct_611_catmull"] = np32(ref.correct_motion(movie, field1.clone(), px))
        fast_in = field1.clone()
        out["correct_fast"] = np32(ref.correct_motion_fast(movie, fast_in))
        out["correct_fast_field_after"] = np32(fast_in)
        out["correct_slow"] = np32(ref.correct_motion_slow(movie, field.clone()))
        new = ref.correct_motion.__globals__["CubicCatmullRomGrid3d"].from_grid_data(smooth_field(6, (3, 4, 4), 0.7))
        base = ref.correct_motion.__globals__["CubicCatmullRomGrid3d"].from_grid_data(field.clone())
        out["two_grids_new"] = np32(new.data)
        out["correct_two_grids"] = np32(
            ref.correct_motion_two_grids(movie, new, base, px, grad=False)
        )
        # pixel shifts of one frame
        lattice = ref.deformation_field_utils.evaluate_deformation_field_at_t(field, 0.4, (40, 40), "bspline")
        out["lattice_t04_bspline"] = np32(lattice)
        grid = deps.coordinate_grid((96, 96))
        out["pixel_shifts"] = np32(ref.get_pixel_shifts(movie[0], px, lattice, grid))
        # spline evaluation
        g = torch.Generator().manual_seed(9)
        tyx = torch.rand((257, 3), generator=g)
        tyx[0] = 0.0
        tyx[1] = 1.0
        out["tyx"] = np32(tyx)
        out["eval_catmull"] = np32(ref.evaluate_deformation_field(field, tyx))
        out["eval_bspline"] = np32(ref.evaluate_deformation_field(field, tyx, grid_type="bspline"))
        out["eval_611_catmull"] = np32(ref.evaluate_deformation_field(field1, tyx))
        out["resample_to_534"] = np32(ref.deformation_field_utils.resample_deformation_field(field, (5, 3, 4)))

        # spline optimiser
        init = torch.as_tensor(out["xc_full_mean_except_current"]).clone()
        for name, kw in {
            "adam_catmull_mse": dict(optimizer_type="adam", grid_type="catmull_rom", loss_type="mse"),
            "adam_bspline_mse": dict(optimizer_type="adam", grid_type="bspline", loss_type="mse"),
            "sgd_bspline_ncc": dict(optimizer_type="sgd", grid_type="bspline", loss_type="ncc"),
            "rmsprop_catmull_cc": dict(optimizer_type="rmsprop", grid_type="catmull_rom", loss_type="cc",
                                       optimizer_kwargs={"lr": 0.001}),
            "lbfgs_bspline_mse": dict(optimizer_type="lbfgs", grid_type="bspline", loss_type="mse"),
        }.items():
            random.seed(1234)
            res, traj = ref.estimate_local_motion(
                movie.clone(), px, (32, 32), (3, 3, 3), init.clone(), n_iterations=6, frequency_range=fr,
                return_trajectory=True, **kw,
            )
            out[f"local_{name}"] = np32(res)
            out[f"local_{name}_losses"] = np.asarray([c.loss for c in traj.checkpoints])
        random.seed(77)
        res = ref.estimate_local_motion(
            movie.clone(), px, (48, 48), (2, 2, 2), None, n_iterations=4, frequency_range=fr
        )
        out["local_noinit_p48"] = np32(res)
    np.savez_compressed(os.path.join(OUT, "small.npz"), **out)
    print("small.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def case_eviction(ref):
    """T=60 > 50 cache entries: the irregular Q1 regime (cache eviction)."""
    out = {}
    px = 1.0
    movie, _ = synthetic_movie(60, 64, 64, seed=21, noise=0.7, drift=2.5, local=0.0, sigma_f=0.09)
    out["movie"] = np32(movie.half().float())  # values exactly representable in fp16 -> smaller file
    movie = torch.as_tensor(out["movie"])
    fr = (100.0, 5.0)
    out["frequency_range"] = np.asarray(fr)
    with verbatim.quiet():
        for strat in ("mean_except_current", "middle_frame"):
            f, _ = ref.estimate_motion_cross_correlation_patches(
                movie.clone(), px, reference_strategy=strat, patch_sidelength=32, frequency_range=fr,
                temporal_smoothing=False, outlier_rejection=False,
            )
            out[f"xc_raw_{strat}"] = np32(f)
    out["movie"] = out["movie"].astype(np.float16)
    np.savez_compressed(os.path.join(OUT, "eviction.npz"), **out)
    print("eviction.npz done")


def case_c1(ref):
    """BASELINE config 1: 10x512x512, known integer global shifts, patch 128 (6x6)."""
    out = {}
    px = 1.0
    movie, walk = synthetic_movie(10, 512, 512, seed=0, noise=1.0, drift=6.0, integer_shifts=True, sigma_f=0.08)
    out["seed"] = np.int64(0)
    out["true_shifts"] = np32(walk)
    with verbatim.quiet():
        g = ref.estimate_global_motion(movie.clone(), px)
        out["global_field"] = np32(g)
        f, pos = ref.estimate_motion_cross_correlation_patches(movie.clone(), px, patch_sidelength=128)
        out["xc_field"] = np32(f)
        out["xc_positions"] = np32(pos)
        corr = ref.correct_motion(movie, f, px, grid_type="bspline")
        s = corr.sum(dim=0)
        out["corrected_sum_crop"] = np32(s[192:320, 192:320])
        out["corrected_sum_norm"] = np.float64(torch.linalg.norm(s.double()))
        out["corrected_sum_rows"] = np32(s[::64, :])
        random.seed(5)
        res, traj = ref.estimate_local_motion(
            movie.clone(), px, (128, 128), (3, 5, 5), f.clone(), n_iterations=3, grid_type="bspline",
            return_trajectory=True,
        )
        out["local_field"] = np32(res)
        out["local_losses"] = np.asarray([c.loss for c in traj.checkpoints])
    np.savez_compressed(os.path.join(OUT, "c1.npz"), **out)
    print("c1.npz done")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref = verbatim.load()
    import torch_motion_correction.deformation_field_utils  # noqa: F401  (attribute access below)

    case_small(ref)
    case_eviction(ref)
    case_c1(ref)


if __name__ == "__main__":
    main()
