"""Movie files either side of the hot path: MRC2014 stacks in, aligned sums out.

The reference's example script reads movies with third-party readers (``mrcfile``, ``eerfile``, ``tifffile``:
``examples/ttMotion.py:1-4,40-62,357``), none of which ship with the package.  This module covers the one container that
needs no codec -- MRC2014 (CCP-EM, Cheng et al. 2015: a 1024-byte header, an optional extended header, raw voxels, x
fastest) -- and hands the frames over in their FILE data type: int8 / uint8 / int16 / uint16 / float16 movies cross PCIe
at 1-2 bytes per pixel and are converted on the device (``tmc_convert_stack`` behind ``motion_correct_many`` /
``prepare_movie``).  EER (Falcon electron-event lists) and compressed TIFF need their decoders and stay out of scope.

PARITY UNPINNED against ``mrcfile`` (absent here); the layout below is the published one and is pinned by
``tests/test_movie_io.py`` on hand-built headers."""

from __future__ import annotations

import os
from pathlib import Path
from typing import Iterable, Iterator, Union

import numpy as np
import torch

_HEADER_BYTES = 1024
#: MRC mode -> numpy dtype of the voxels (little-endian files; byte-swapped when the machine stamp says big-endian)
_MODE_DTYPES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}
_DTYPE_MODES = {np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.float32): 2, np.dtype(np.uint16): 6,
                np.dtype(np.float16): 12}


class MrcHeader:
    """The fields of the 1024-byte MRC2014 header this package uses."""

    def __init__(self, nx: int, ny: int, nz: int, mode: int, cella=(0.0, 0.0, 0.0), mx: int = 0, my: int = 0, mz: int = 0,
                 nsymbt: int = 0, big_endian: bool = False, unsigned_bytes: bool = False):
        self.nx, self.ny, self.nz, self.mode = nx, ny, nz, mode
        self.cella = tuple(float(v) for v in cella)
        self.mx, self.my, self.mz = mx or nx, my or ny, mz or nz
        self.nsymbt = nsymbt
        self.big_endian = big_endian
        self.unsigned_bytes = unsigned_bytes

    @property
    def pixel_spacing(self) -> float:
        """Angstrom per pixel along x (cell length / grid sampling), 0 if the file does not say."""
        return self.cella[0] / self.mx if self.mx and self.cella[0] > 0 else 0.0

    @property
    def dtype(self) -> np.dtype:
        if self.mode not in _MODE_DTYPES:
            raise ValueError(f"MRC mode {self.mode} is not a real-valued image type this package reads "
                             f"(supported: {sorted(_MODE_DTYPES)})")
        dt = np.dtype(np.uint8 if (self.mode == 0 and self.unsigned_bytes) else _MODE_DTYPES[self.mode])
        return dt.newbyteorder(">") if self.big_endian and dt.itemsize > 1 else dt

    @property
    def data_offset(self) -> int:
        return _HEADER_BYTES + self.nsymbt


def read_mrc_header(path: Union[str, Path]) -> MrcHeader:
    """Parse the fixed header: words 1-4 nx ny nz mode, 8-10 mx my mz, 11-13 cell lengths, 24 nsymbt, bytes 208-211 'MAP ',
    212-215 machine stamp (0x44 little endian, 0x11 big endian).  Mode 0 is signed unless the IMOD flag (bytes 152-159:
    stamp 1146047817, bit 0 of the flags word) says unsigned."""
    with open(path, "rb") as f:
        raw = f.read(_HEADER_BYTES)
    if len(raw) < _HEADER_BYTES:
        raise ValueError(f"{path}: shorter than an MRC header")
    stamp = raw[212]
    big = stamp == 0x11
    order = ">" if big else "<"
    i4 = np.frombuffer(raw, dtype=np.dtype(order + "i4"), count=56)
    f4 = np.frombuffer(raw, dtype=np.dtype(order + "f4"), count=56)
    nx, ny, nz, mode = (int(v) for v in i4[:4])
    if raw[208:211] != b"MAP" and not (0 < nx < 1 << 20 and 0 < ny < 1 << 20 and 0 < nz < 1 << 20 and 0 <= mode <= 101):
        raise ValueError(f"{path}: not an MRC file")
    imod_unsigned = int(i4[38]) == 1146047817 and not (int(i4[39]) & 1)  # IMOD: bit 0 set = signed bytes
    return MrcHeader(nx, ny, nz, mode, cella=f4[10:13], mx=int(i4[7]), my=int(i4[8]), mz=int(i4[9]), nsymbt=max(int(i4[23]), 0),
                     big_endian=big, unsigned_bytes=imod_unsigned)


def read_mrc(path: Union[str, Path], pinned: bool = True, frames: slice | None = None):
    """``(movie, header)``: the (nz, ny, nx) stack as a HOST tensor in the file's data type (int8 / uint8 / int16 / uint16 /
    float16 / float32), page-locked by default so that ``motion_correct_many`` can copy it asynchronously.

    The voxels are read straight into the (pinned) destination buffer; big-endian files are byte-swapped on the way."""
    hdr = read_mrc_header(path)
    dt = hdr.dtype
    first, last, step = (frames or slice(None)).indices(hdr.nz)
    if step != 1:
        raise ValueError("read_mrc: frames must be a contiguous slice")
    n = max(last - first, 0)
    plane = hdr.ny * hdr.nx
    need = hdr.data_offset + (first + n) * plane * dt.itemsize
    if os.path.getsize(path) < need:
        raise ValueError(f"{path}: truncated ({os.path.getsize(path)} bytes, header promises {need})")
    native = dt.newbyteorder("=")
    torch_dtype = {np.dtype(np.int8): torch.int8, np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
                   np.dtype(np.uint16): torch.uint16, np.dtype(np.float16): torch.float16,
                   np.dtype(np.float32): torch.float32}[native]
    out = torch.empty((n, hdr.ny, hdr.nx), dtype=torch_dtype, pin_memory=bool(pinned and torch.cuda.is_available()))
    view = out.numpy() if torch_dtype != torch.uint16 else out.view(torch.int16).numpy().view(np.uint16)
    with open(path, "rb") as f:
        f.seek(hdr.data_offset + first * plane * dt.itemsize)
        got = f.readinto(memoryview(view.reshape(-1)).cast("B"))
    if got != n * plane * dt.itemsize:
        raise ValueError(f"{path}: short read")
    if dt.byteorder == ">":
        view.byteswap(inplace=True)
    return out, hdr


def write_mrc(path: Union[str, Path], data, pixel_spacing: float = 0.0, overwrite: bool = True) -> None:
    """Write a (ny, nx) image or (nz, ny, nx) stack (tensor or array; int8 / int16 / uint16 / float16 / float32; anything
    else is stored as float32) as a little-endian MRC2014 file with density statistics and the cell set from
    ``pixel_spacing`` (Angstrom per pixel)."""
    arr = data.detach().cpu() if isinstance(data, torch.Tensor) else data
    if isinstance(arr, torch.Tensor):
        arr = arr.view(torch.int16).numpy().view(np.uint16) if arr.dtype == torch.uint16 else arr.numpy()
    arr = np.asarray(arr)
    if arr.ndim == 2:
        arr = arr[None]
    if arr.ndim != 3:
        raise ValueError(f"write_mrc: expected a (ny, nx) image or (nz, ny, nx) stack, got shape {arr.shape}")
    if arr.dtype not in _DTYPE_MODES:
        arr = arr.astype(np.float32)
    arr = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<"), copy=False))
    nz, ny, nx = arr.shape
    if not overwrite and os.path.exists(path):
        raise FileExistsError(path)
    head = np.zeros(256, dtype="<i4")
    fview = head.view("<f4")
    head[0:3] = (nx, ny, nz)
    head[3] = _DTYPE_MODES[arr.dtype.newbyteorder("=")]
    head[7:10] = (nx, ny, nz)
    fview[10:13] = (nx * pixel_spacing, ny * pixel_spacing, nz * pixel_spacing)
    fview[13:16] = 90.0
    head[16:19] = (1, 2, 3)
    stats = arr.astype(np.float64) if arr.size else np.zeros(1)
    fview[19:22] = (stats.min(), stats.max(), stats.mean())
    fview[54] = stats.std()
    head[22] = 1 if nz == 1 else 0  # space group: 1 = volume-like single image, 0 = image stack
    head[27] = 20140  # NVERSION
    raw = bytearray(head.tobytes())
    raw[208:212] = b"MAP "
    raw[212:216] = bytes([0x44, 0x44, 0x00, 0x00])
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(raw)
        f.write(memoryview(arr.reshape(-1)).cast("B"))


def mrc_movies(paths: Iterable[Union[str, Path]], pinned: bool = True) -> Iterator[torch.Tensor]:
    """The movies of ``paths`` one after the other as pinned host tensors in their file data type: the input
    ``motion_correct_many`` pipelines (disk read of movie i+1 overlaps the device work on movie i through its prefetch)."""
    for p in paths:
        yield read_mrc(p, pinned=pinned)[0]
