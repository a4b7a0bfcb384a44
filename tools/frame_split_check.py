#!/usr/bin/env python
"""NCCL check + timing of the frame-split path (one movie over N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/frame_split_check.py [--t 40 --size 4096 --patch 1024]

Every rank generates the same movie, keeps its frame block, runs motion_correct_frame_split and
compares with the single-GPU pipeline run on rank 0.  Prints one JSON line on rank 0."""

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=16)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--patch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import torch_motion_correction_b200 as tmc
    from torch_motion_correction_b200.distributed import frame_range, motion_correct_frame_split

    movie, _ = bench.synthetic_movie_gpu(args.t, args.size, args.size, 7, dev)
    f0, f1 = frame_range(args.t, rank, world)
    local_frames = movie[f0:f1].contiguous()
    px = 0.83

    def step():
        return motion_correct_frame_split(local_frames, px, f0, args.t, patch_sidelength=args.patch)

    total, field = step()
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        total, field = step()
    e.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    want_total, want_field = tmc.motion_correct(movie, px, patch_sidelength=args.patch)
    torch.cuda.synchronize()
    s.record()
    for _ in range(args.steps):
        tmc.motion_correct(movie, px, patch_sidelength=args.patch)
    e.record()
    torch.cuda.synchronize()
    single_ms = s.elapsed_time(e) / args.steps
    err_field = float((field - want_field).abs().max())
    err_sum = float(torch.linalg.norm(total - want_total) / torch.linalg.norm(want_total))
    if rank == 0:
        print(json.dumps({
            "check": "frame_split_vs_single_gpu", "world": world, "movie": [args.t, args.size, args.size], "patch": args.patch,
            "max_abs_field_diff_angstrom": err_field, "rel_l2_sum_diff": err_sum, "frame_split_ms_per_movie": float(ms),
            "single_gpu_ms_per_movie": single_ms, "ok": err_field <= 1e-3 and err_sum <= 1e-5,
        }))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
