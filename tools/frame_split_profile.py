#!/usr/bin/env python
"""Per-entry-point device time of the frame-split path (rank 0) next to the single-GPU pipeline, C4 by default:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/frame_split_profile.py [--iterations 100]"""
import argparse, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench


def timed(fn, lib, reps=2):
    fn()
    torch.cuda.synchronize()
    lib.TIMING = {}
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    timing, lib.TIMING = lib.TIMING, None
    per = {k: round(sum(a.elapsed_time(b) for a, b in v) / reps, 3) for k, v in timing.items()}
    return s.elapsed_time(e) / reps, wall, dict(sorted(per.items(), key=lambda kv: -kv[1])[:12])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=100)
    ap.add_argument("--workload", default="c4")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import torch_motion_correction_b200 as tmc
    from torch_motion_correction_b200 import _lib
    from torch_motion_correction_b200.distributed import frame_range, motion_correct_frame_split

    cfg = bench.WORKLOADS[args.workload]
    t, h, w, px, p = cfg["t"], cfg["h"], cfg["w"], cfg["pixel_spacing"], cfg["patch"]
    f0, f1 = frame_range(t, rank, world)
    local_frames, _ = bench.synthetic_frames_gpu(t, h, w, 4040, dev, f0, f1)
    kw = dict(patch_sidelength=p, n_iterations=args.iterations, deformation_field_resolution=cfg["resolution"])

    def split():
        random.seed(99)
        return motion_correct_frame_split(local_frames, px, f0, t, **kw)

    out = {}
    dist.barrier()
    ms, wall, per = timed(split, _lib)
    out["split"] = {"device_ms": round(ms, 2), "wall_ms": round(wall, 2), "entries": per}
    dist.barrier()
    if rank == 0:
        del local_frames
        torch.cuda.empty_cache()
        movie, _ = bench.synthetic_frames_gpu(t, h, w, 4040, dev)

        def single():
            random.seed(99)
            return tmc.motion_correct(movie, px, **kw)

        ms, wall, per = timed(single, _lib)
        out["single"] = {"device_ms": round(ms, 2), "wall_ms": round(wall, 2), "entries": per}
        print(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
