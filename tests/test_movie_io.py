"""MRC2014 movie files (torch_motion_correction_b200.movie_io): header layout pinned on hand-built bytes, round trips of
every supported voxel type, extended headers, big-endian files, error cases.  CPU only."""

import struct

import numpy as np
import pytest
import torch

from torch_motion_correction_b200 import movie_io


def hand_built_mrc(nx, ny, nz, mode, payload: bytes, cell=(0.0, 0.0, 0.0), ext: bytes = b"", big=False, imod_flags=None):
    """A file assembled field by field from the published MRC2014 byte offsets (independent of write_mrc)."""
    e = ">" if big else "<"
    h = bytearray(1024)
    struct.pack_into(e + "4i", h, 0, nx, ny, nz, mode)          # words 1-4: columns, rows, sections, mode
    struct.pack_into(e + "3i", h, 28, nx, ny, nz)               # words 8-10: grid sampling mx my mz
    struct.pack_into(e + "3f", h, 40, *cell)                    # words 11-13: cell lengths (Angstrom)
    struct.pack_into(e + "3f", h, 52, 90.0, 90.0, 90.0)         # words 14-16: cell angles
    struct.pack_into(e + "3i", h, 64, 1, 2, 3)                  # words 17-19: axis order
    struct.pack_into(e + "i", h, 92, len(ext))                  # word 24: bytes of extended header
    if imod_flags is not None:
        struct.pack_into(e + "2i", h, 152, 1146047817, imod_flags)
    h[208:212] = b"MAP "
    h[212:216] = bytes([0x11, 0x11, 0, 0]) if big else bytes([0x44, 0x44, 0, 0])
    return bytes(h) + ext + payload


def test_header_fields_of_a_hand_built_file(tmp_path):
    data = np.arange(2 * 3 * 4, dtype="<u2").reshape(2, 3, 4)
    path = tmp_path / "stack.mrc"
    path.write_bytes(hand_built_mrc(4, 3, 2, 6, data.tobytes(), cell=(4 * 0.83, 3 * 0.83, 2 * 0.83), ext=b"\x07" * 128))
    hdr = movie_io.read_mrc_header(path)
    assert (hdr.nx, hdr.ny, hdr.nz, hdr.mode, hdr.nsymbt) == (4, 3, 2, 6, 128)
    assert abs(hdr.pixel_spacing - 0.83) < 1e-6
    assert hdr.data_offset == 1024 + 128
    movie, _ = movie_io.read_mrc(path, pinned=False)
    assert movie.dtype == torch.uint16 and tuple(movie.shape) == (2, 3, 4)
    assert np.array_equal(movie.view(torch.int16).numpy().view(np.uint16), data)
    first, _ = movie_io.read_mrc(path, pinned=False, frames=slice(1, 2))
    assert np.array_equal(first.view(torch.int16).numpy().view(np.uint16), data[1:2])


@pytest.mark.parametrize("mode,dtype", [(0, np.int8), (1, np.int16), (2, np.float32), (6, np.uint16), (12, np.float16)])
@pytest.mark.parametrize("big", [False, True])
def test_hand_built_files_of_every_mode_and_byte_order(tmp_path, mode, dtype, big):
    rng = np.random.default_rng(mode)
    data = (rng.normal(size=(3, 5, 6)) * 50).astype(dtype)
    stored = data.astype(np.dtype(dtype).newbyteorder(">" if big else "<"))
    path = tmp_path / "m.mrc"
    path.write_bytes(hand_built_mrc(6, 5, 3, mode, stored.tobytes(), big=big))
    movie, hdr = movie_io.read_mrc(path, pinned=False)
    assert hdr.big_endian == big
    got = movie.view(torch.int16).numpy().view(np.uint16) if movie.dtype == torch.uint16 else movie.numpy()
    assert got.dtype == np.dtype(dtype) and np.array_equal(got, data)


def test_imod_unsigned_bytes(tmp_path):
    data = np.array([[[0, 127, 128, 255]]], dtype=np.uint8)
    path = tmp_path / "u8.mrc"
    path.write_bytes(hand_built_mrc(4, 1, 1, 0, data.tobytes(), imod_flags=0))  # stamp present, bit 0 clear: unsigned
    movie, _ = movie_io.read_mrc(path, pinned=False)
    assert movie.dtype == torch.uint8 and movie.flatten().tolist() == [0, 127, 128, 255]
    path.write_bytes(hand_built_mrc(4, 1, 1, 0, data.tobytes(), imod_flags=1))  # bit 0 set: signed
    movie, _ = movie_io.read_mrc(path, pinned=False)
    assert movie.dtype == torch.int8 and movie.flatten().tolist() == [0, 127, -128, -1]


@pytest.mark.parametrize("dtype", [torch.int8, torch.int16, torch.uint16, torch.float16, torch.float32])
def test_write_then_read(tmp_path, dtype):
    g = torch.Generator().manual_seed(5)
    data = (torch.randn((4, 7, 9), generator=g) * 40).to(torch.float32)
    data = data.abs().to(dtype) if dtype == torch.uint16 else data.to(dtype)
    path = tmp_path / "sub" / "w.mrc"
    movie_io.write_mrc(path, data, pixel_spacing=1.25)
    back, hdr = movie_io.read_mrc(path, pinned=False)
    assert back.dtype == dtype and torch.equal(back.view(torch.int16) if dtype == torch.uint16 else back,
                                               data.view(torch.int16) if dtype == torch.uint16 else data)
    assert abs(hdr.pixel_spacing - 1.25) < 1e-6 and (hdr.mx, hdr.my, hdr.mz) == (9, 7, 4)
    raw = path.read_bytes()
    assert raw[208:212] == b"MAP " and raw[212] == 0x44 and len(raw) == 1024 + data.numel() * data.element_size()
    dmin, dmax, dmean = struct.unpack_from("<3f", raw, 76)
    ref = data.view(torch.int16).numpy().view(np.uint16).astype(np.float64) if dtype == torch.uint16 else data.double().numpy()
    assert abs(dmin - ref.min()) < 1e-3 and abs(dmax - ref.max()) < 1e-3 and abs(dmean - ref.mean()) < 1e-3


def test_single_image_and_other_types_are_stored_as_float32(tmp_path):
    img = torch.arange(12, dtype=torch.float64).reshape(3, 4)
    path = tmp_path / "sum.mrc"
    movie_io.write_mrc(path, img)
    back, hdr = movie_io.read_mrc(path, pinned=False)
    assert hdr.mode == 2 and tuple(back.shape) == (1, 3, 4) and torch.equal(back[0], img.float())
    with pytest.raises(FileExistsError):
        movie_io.write_mrc(path, img, overwrite=False)


def test_errors(tmp_path):
    path = tmp_path / "bad.mrc"
    path.write_bytes(b"\x00" * 100)
    with pytest.raises(ValueError, match="shorter than an MRC header"):
        movie_io.read_mrc_header(path)
    path.write_bytes(hand_built_mrc(4, 4, 4, 2, b"\x00" * 10))
    with pytest.raises(ValueError, match="truncated"):
        movie_io.read_mrc(path, pinned=False)
    path.write_bytes(hand_built_mrc(4, 4, 1, 4, b"\x00" * 128))  # complex64 transform: not an image stack
    with pytest.raises(ValueError, match="mode 4"):
        movie_io.read_mrc(path, pinned=False)
    with pytest.raises(ValueError, match="contiguous"):
        path.write_bytes(hand_built_mrc(2, 2, 4, 2, b"\x00" * 64))
        movie_io.read_mrc(path, pinned=False, frames=slice(0, 4, 2))


def test_mrc_movies_generator(tmp_path):
    paths = []
    for i in range(3):
        p = tmp_path / f"m{i}.mrc"
        movie_io.write_mrc(p, torch.full((2, 4, 4), i, dtype=torch.int16))
        paths.append(p)
    movies = list(movie_io.mrc_movies(paths, pinned=False))
    assert [int(m[0, 0, 0]) for m in movies] == [0, 1, 2] and all(m.dtype == torch.int16 for m in movies)
