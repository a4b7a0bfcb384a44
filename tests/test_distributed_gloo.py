"""world_size-2 gloo run (CPU) of the frame-split bookkeeping: frame blocks, padded all-gather,
moment all-reduce.  The kernels themselves are covered by the -m gpu tests."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_motion_correction_b200.distributed import _reshard_frames_to_patches, all_gather_frames, frame_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f0, f1 = frame_range(total_frames, rank, world)
        # every "frame" carries its global index: gathering must restore frame order
        local = torch.arange(f0, f1, dtype=torch.float32)[:, None].repeat(1, 3)
        full = all_gather_frames(local, total_frames)
        assert full.shape == (total_frames, 3)
        assert torch.equal(full[:, 0], torch.arange(total_frames, dtype=torch.float32))
        # moments: sum / sum of squares / count all-reduce == whole-movie statistics
        g = torch.Generator().manual_seed(5)
        movie = torch.randn((total_frames, 8, 8), generator=g, dtype=torch.float64) * 2 + 3
        mine = movie[f0:f1]
        moments = torch.stack([mine.sum(), (mine * mine).sum(), torch.tensor(float(mine.numel()), dtype=torch.float64)])
        dist.all_reduce(moments, op=dist.ReduceOp.SUM)
        mean = moments[0] / moments[2]
        std = torch.sqrt((moments[1] - moments[0] * mean) / (moments[2] - 1))
        assert abs(float(mean - movie.mean())) < 1e-12 and abs(float(std - movie.std())) < 1e-10
        # frame-sum all-reduce
        part = mine.sum(dim=0)
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        assert torch.allclose(part, movie.sum(dim=0))
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_frames", [7, 10])
def test_frame_split_bookkeeping_world2(tmp_path, total_frames):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total_frames, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _reshard_worker(rank, world, port, t, g, out_dir):
    """The frame-split optimiser's re-sharding: spectra computed where the frames live, (t_local, G, words), gathered to
    (T, G, words) and cut into this rank's share of the patches with ALL frames."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        words = 5
        full = torch.arange(t * g * words, dtype=torch.float32).reshape(t, g, words)
        f0, f1 = frame_range(t, rank, world)
        gathered = all_gather_frames(full[f0:f1].contiguous(), t)
        assert torch.equal(gathered, full)
        g0, g1 = frame_range(g, rank, world)
        mine = gathered[:, g0:g1].permute(1, 0, 2).contiguous()  # (G_r, T, words)
        assert mine.shape == (g1 - g0, t, words)
        # the product's exchange (all-to-all of the planes each peer owns) delivers exactly that block
        tp = 2 * ((t + 1) // 2)
        sub = torch.zeros((max(g1 - g0, 1), tp, words))
        _reshard_frames_to_patches(full[f0:f1].contiguous(), sub, t, g, words, rank, world, None)
        assert torch.equal(sub[: g1 - g0, :t], mine) and float(sub[:, t:].abs().sum()) == 0.0
        # every patch is owned by exactly one rank
        owned = torch.zeros(g)
        owned[g0:g1] = 1
        dist.all_reduce(owned)
        assert torch.equal(owned, torch.ones(g))
        # the coefficient gradient is a sum over patches: partial sums all-reduce to the whole
        part = mine.sum(dim=(0, 1))
        dist.all_reduce(part)
        assert torch.allclose(part, full.sum(dim=(0, 1)))
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("t,g", [(7, 9), (10, 3)])
def test_frame_split_patch_resharding_world2(tmp_path, t, g):
    world = 2
    mp.spawn(_reshard_worker, args=(world, _free_port(), t, g, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
