"""Patch-grid geometry (mirror of the reference's ``patch_grid`` package; geometry only).

The reference's eager / lazy patch *extraction* (``patch_grid``, ``patch_grid_lazy``) is an
implementation detail of its estimators: here patches are never materialised -- the FFT row
kernel reads them straight out of the movie (``csrc/fourier.cu``).  The centre placement is
reproduced exactly because the centres are returned to the user."""

from ._patch_grid_centers import patch_centers_1d, patch_grid_centers

__all__ = ["patch_grid_centers", "patch_centers_1d"]
