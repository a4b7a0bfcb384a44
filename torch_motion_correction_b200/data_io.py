"""Deformation-field CSV I/O (mirror of the reference's ``data_io.py``: columns t,h,w,y_shift,x_shift)."""

from __future__ import annotations

from pathlib import Path
from typing import Union

import numpy as np
import torch


def write_deformation_field_to_csv(deformation_field: torch.Tensor, output_path: Union[str, Path]) -> None:
    """(2, t, h, w) field -> CSV rows ``t,h,w,y_shift,x_shift`` (reference data_io.py:10-73)."""
    _, t, h, w = deformation_field.shape
    field = deformation_field.detach().cpu().to(torch.float64).numpy()
    tt, hh, ww = np.meshgrid(np.arange(t), np.arange(h), np.arange(w), indexing="ij")
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    with open(output_path, "w") as f:
        f.write("t,h,w,y_shift,x_shift\n")
        for a, b, c, y, x in zip(tt.ravel(), hh.ravel(), ww.ravel(), field[0].ravel(), field[1].ravel()):
            f.write(f"{a},{b},{c},{float(y)!r},{float(x)!r}\n")


def read_deformation_field_from_csv(csv_path: Union[str, Path], device: torch.device = None) -> torch.Tensor:
    """CSV -> (2, t, h, w) float32 field (reference data_io.py:76-141)."""
    if device is None:
        device = torch.device("cpu")
    rows = np.genfromtxt(csv_path, delimiter=",", names=True, dtype=None, encoding="utf-8")
    rows = np.atleast_1d(rows)
    ut, uh, uw = (np.unique(rows[k]) for k in ("t", "h", "w"))
    field = np.zeros((2, len(ut), len(uh), len(uw)), dtype=np.float32)
    ti, hi, wi = (np.searchsorted(u, rows[k]) for u, k in ((ut, "t"), (uh, "h"), (uw, "w")))
    field[0, ti, hi, wi] = rows["y_shift"]
    field[1, ti, hi, wi] = rows["x_shift"]
    return torch.as_tensor(field, device=device)
