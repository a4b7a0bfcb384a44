"""Movie preparation (detector-native pixel types, gain, hot pixels, per-frame mean removal) against the oracle's
restatement of the reference example's NumPy pre-processing (examples/ttMotion.py:90-202), and movies arriving from the
host in their native type through motion_correct_many."""

import numpy as np
import pytest
import torch

import torch_motion_correction_b200 as tmc
from oracle import reference_path as rp

pytestmark = pytest.mark.gpu

DTYPES = {"uint8": torch.uint8, "int8": torch.int8, "uint16": torch.uint16, "int16": torch.int16, "float16": torch.float16,
          "float32": torch.float32}


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def counts_movie(t, h, w, seed, np_dtype):
    rng = np.random.default_rng(seed)
    movie = rng.poisson(30.0, size=(t, h, w)).astype(np.float64)
    movie[0, 0, 0] = 250.0  # hot pixels: a corner, an edge, the interior, a dead one
    movie[1, h // 2, 0] = 240.0
    movie[t - 1, h // 3, w // 3] = 255.0
    movie[2, 5, 7] = 0.0
    return movie.astype(np_dtype)


@pytest.mark.parametrize("name", list(DTYPES))
@pytest.mark.parametrize("shape", [(5, 64, 96), (3, 37, 53)])
def test_prepare_movie_matches_the_example_pre_processing(dev, name, shape):
    t, h, w = shape
    np_dtype = {"uint8": np.uint8, "int8": np.int8, "uint16": np.uint16, "int16": np.int16, "float16": np.float16,
                "float32": np.float32}[name]
    movie = counts_movie(t, h, w, 3, np.float64)
    if name == "int8":  # signed bytes (MRC mode 0): the same counts, centred
        movie = movie - 128.0
    movie = movie.astype(np_dtype)
    rng = np.random.default_rng(1)
    gain = (1.0 + 0.05 * rng.standard_normal((h, w))).astype(np.float32)
    want, n_hot = rp.prepare_movie(movie, gain=gain, hot_pixel_threshold=10.0, zero_frame_means=True)
    assert n_hot >= 3
    host = torch.from_numpy(movie.view(np.uint16) if name == "uint16" else movie)
    if name == "uint16":
        host = host.view(torch.uint16)
    got, count = tmc.prepare_movie(host, gain=torch.from_numpy(gain), hot_pixel_threshold=10.0, zero_frame_means=True, device=dev,
                                   return_hot_pixel_count=True)
    assert got.dtype == torch.float32 and got.shape == (t, h, w)
    assert int(count) == n_hot
    assert float((got.cpu() - torch.from_numpy(want)).abs().max()) <= 2e-4
    # conversion + gain alone are exact
    plain = tmc.prepare_movie(host, gain=torch.from_numpy(gain), device=dev)
    assert torch.equal(plain.cpu(), torch.from_numpy(movie.astype(np.float32) * gain))
    assert torch.equal(tmc.prepare_movie(host, device=dev).cpu(), torch.from_numpy(movie.astype(np.float32)))


def test_native_type_movies_through_the_pipeline(dev):
    """uint16 counts from (pinned) host memory: same sums / fields as the float32 movie, a quarter... half of the PCIe bytes."""
    movies = []
    for s in (1, 2, 3):
        m, _ = rp.synthetic_movie(5, 128, 128, seed=s, noise=0.5, drift=2.0, local=0.3)
        movies.append(torch.round(m * 40 + 400).clamp(0, 65535))
    kwargs = dict(patch_sidelength=64, frequency_range=(80, 5), n_iterations=0)
    as_u16 = [torch.from_numpy(m.numpy().astype(np.uint16)).view(torch.uint16).pin_memory() for m in movies]
    got = [(s.clone(), f.cpu()) for s, f in tmc.motion_correct_many(as_u16, 1.1, device=dev, **kwargs)]
    assert len(got) == 3
    for m, (host_sum, field) in zip(movies, got):
        want_sum, want_field = tmc.motion_correct(m.to(dev), 1.1, **kwargs)
        assert torch.equal(field, want_field.cpu())
        assert float(torch.linalg.norm(host_sum - want_sum.cpu()) / torch.linalg.norm(want_sum.cpu())) <= 1e-6


def test_mrc_files_through_the_pipeline(dev, tmp_path):
    """Movies stored as MRC2014 stacks of unsigned 16-bit and signed 8-bit counts: read in the file's type into pinned
    memory, aligned through motion_correct_many, sums written back as MRC -- against the float32 in-memory route."""
    movies, paths = [], []
    for i, (s, dtype) in enumerate(((1, torch.uint16), (2, torch.int8))):
        m, _ = rp.synthetic_movie(5, 128, 128, seed=s, noise=0.5, drift=2.0, local=0.3)
        m = torch.round(m * 40 + 400).clamp(0, 65535) if dtype == torch.uint16 else torch.round(m * 12).clamp(-128, 127)
        movies.append(m)
        paths.append(tmp_path / f"movie{i}.mrc")
        stored = torch.from_numpy(m.numpy().astype(np.uint16)).view(torch.uint16) if dtype == torch.uint16 else m.to(torch.int8)
        tmc.write_mrc(paths[-1], stored, pixel_spacing=1.1)
    assert tmc.read_mrc_header(paths[0]).mode == 6 and tmc.read_mrc_header(paths[1]).mode == 0
    kwargs = dict(patch_sidelength=64, frequency_range=(80, 5), n_iterations=0)
    px = tmc.read_mrc_header(paths[0]).pixel_spacing
    got = [(s.clone(), f.cpu()) for s, f in tmc.motion_correct_many(tmc.mrc_movies(paths), px, device=dev, **kwargs)]
    for i, (m, (host_sum, field)) in enumerate(zip(movies, got)):
        want_sum, want_field = tmc.motion_correct(m.to(dev), px, **kwargs)
        assert torch.equal(field, want_field.cpu())
        assert float(torch.linalg.norm(host_sum - want_sum.cpu()) / torch.linalg.norm(want_sum.cpu())) <= 1e-6
        tmc.write_mrc(tmp_path / f"sum{i}.mrc", host_sum, pixel_spacing=px)
        back, hdr = tmc.read_mrc(tmp_path / f"sum{i}.mrc", pinned=False)
        assert torch.equal(back[0], host_sum) and abs(hdr.pixel_spacing - px) < 1e-6
