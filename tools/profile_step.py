#!/usr/bin/env python
"""One motion_correct step of a bench workload bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off` (see /opt/skills/guides/B200_PROFILING.md and profiles/README.md)."""

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import torch_motion_correction_b200 as tmc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--iterations", type=int, default=0)
    ap.add_argument("--only", default="all", choices=["all", "warp", "xc", "global"])
    args = ap.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    dev = torch.device("cuda:0")
    movie, _ = bench.synthetic_movie_gpu(cfg["t"], cfg["h"], cfg["w"], 1000, dev)
    px, p = cfg["pixel_spacing"], cfg["patch"]

    def step():
        if args.only == "warp":
            field = torch.zeros((2, 3, 5, 5), device=dev)
            return tmc.correct_motion_sum(movie, field, px, grid_type="bspline")
        if args.only == "global":
            return tmc.estimate_global_motion(movie, px)
        if args.only == "xc":
            return tmc.estimate_motion_cross_correlation_patches(movie, px, patch_sidelength=p)
        return tmc.motion_correct(movie, px, patch_sidelength=p, deformation_field_resolution=cfg["resolution"],
                                  n_iterations=args.iterations)

    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled step done")


if __name__ == "__main__":
    main()
