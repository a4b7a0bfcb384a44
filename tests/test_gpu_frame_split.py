"""Frame-split (one movie over several ranks) vs the single-GPU pipeline.  Two gloo ranks share
cuda:0 here (NCCL needs one GPU per rank; the NCCL path is exercised by tools/frame_split_check.py
under `gpurun --gpus 2`)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _movie():
    from oracle import reference_path as rp

    movie, _ = rp.synthetic_movie(7, 128, 128, seed=9, noise=0.6, drift=3.0, local=0.5, sigma_f=0.08)
    return movie


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch_motion_correction_b200 as tmc
        from torch_motion_correction_b200.distributed import frame_range, motion_correct_frame_split

        dev = torch.device("cuda:0")
        movie = _movie()
        t = movie.shape[0]
        f0, f1 = frame_range(t, rank, world)
        total, field = motion_correct_frame_split(
            movie[f0:f1].to(dev), 1.1, f0, t, patch_sidelength=64, frequency_range=(80, 5)
        )
        want_total, want_field = tmc.motion_correct(movie.to(dev), 1.1, patch_sidelength=64, frequency_range=(80, 5))
        assert float((field - want_field).abs().max()) <= 1e-4, float((field - want_field).abs().max())
        rel = float(torch.linalg.norm(total - want_total) / torch.linalg.norm(want_total))
        assert rel <= 1e-5, rel
        # with the spline optimiser: Sigma and the coefficient gradient all-reduced every iteration
        import random

        kw = dict(patch_sidelength=64, frequency_range=(80, 5), n_iterations=6, deformation_field_resolution=(3, 3, 3))
        random.seed(11)  # rank 0 draws the mini-batches and broadcasts them
        total, field = motion_correct_frame_split(movie[f0:f1].to(dev), 1.1, f0, t, **kw)
        random.seed(11)
        want_total, want_field = tmc.motion_correct(movie.to(dev), 1.1, **kw)
        assert field.shape == (2, 3, 3, 3)
        assert float((field - want_field).abs().max()) <= 1e-4, float((field - want_field).abs().max())
        assert float(want_field.abs().max()) > 1e-3
        rel = float(torch.linalg.norm(total - want_total) / torch.linalg.norm(want_total))
        assert rel <= 1e-5, rel
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_frame_split_matches_single_gpu_world2(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))


def test_frame_split_world1_degenerates_to_single_gpu():
    import torch_motion_correction_b200 as tmc
    from torch_motion_correction_b200.distributed import motion_correct_frame_split

    dev = torch.device("cuda:0")
    movie = _movie().to(dev)
    total, field = motion_correct_frame_split(movie, 1.1, 0, movie.shape[0], patch_sidelength=64, frequency_range=(80, 5))
    want_total, want_field = tmc.motion_correct(movie, 1.1, patch_sidelength=64, frequency_range=(80, 5))
    assert float((field - want_field).abs().max()) <= 1e-5
    assert float(torch.linalg.norm(total - want_total) / torch.linalg.norm(want_total)) <= 1e-6
    # the split optimiser (two halves around the all-reduce) against the one-launch-per-iteration kernels
    import random

    from torch_motion_correction_b200 import _ops
    from torch_motion_correction_b200.distributed import estimate_local_motion_frame_split

    init = want_field
    random.seed(4)
    want = tmc.estimate_local_motion(movie, 1.1, (64, 64), (3, 3, 3), init, n_iterations=8, frequency_range=(80, 5),
                                     grid_type="bspline")
    random.seed(4)
    got, losses = estimate_local_motion_frame_split(movie, 1.1, (64, 64), (3, 3, 3), init, 0, movie.shape[0], _ops.stack_stats(movie),
                                                    n_iterations=8, frequency_range=(80, 5), grid_type="bspline", return_losses=True)
    assert float((got - want).abs().max()) <= 1e-4, float((got - want).abs().max())
    assert len(losses) == 8 and all(l > 0 for l in losses)


def test_motion_correct_many_matches_single_calls():
    """Pipelined host->device->host processing returns the same sums/fields as one call per movie."""
    import torch_motion_correction_b200 as tmc
    from oracle import reference_path as rp

    dev = torch.device("cuda:0")
    movies = [rp.synthetic_movie(5, 128, 128, seed=s, noise=0.5, drift=2.0, local=0.3)[0].pin_memory() for s in (1, 2, 3)]
    kwargs = dict(patch_sidelength=64, frequency_range=(80, 5), n_iterations=0)
    got = []
    for host_sum, field in tmc.motion_correct_many(movies, 1.1, device=dev, **kwargs):
        torch.cuda.synchronize()
        got.append((host_sum.clone(), field.cpu()))
    assert len(got) == 3
    for m, (host_sum, field) in zip(movies, got):
        want_sum, want_field = tmc.motion_correct(m.to(dev), 1.1, **kwargs)
        assert torch.equal(field, want_field.cpu())
        assert float(torch.linalg.norm(host_sum - want_sum.cpu()) / torch.linalg.norm(want_sum.cpu())) <= 1e-6
