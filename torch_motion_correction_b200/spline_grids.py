"""Learnable uniform cubic spline grids evaluated by the CUDA spline kernels.

Drop-in for the two classes the reference imports from ``torch_cubic_spline_grids``
(``CubicBSplineGrid3d`` / ``CubicCatmullRomGrid3d``; call sites: correct_motion.py:6,188-317,
estimate_motion_optimizer.py:9,122-158; semantics: SURVEY.md Appendix A.1): ``resolution=`` /
``n_channels=`` constructor, ``from_grid_data``, a ``.data`` ``nn.Parameter`` of shape
``(c, n0, n1, n2)``, ``grid(u)`` with ``u (..., 3)`` in [0, 1] -> ``(..., c)``, differentiable
w.r.t. ``.data`` (the transpose scatter of the same 64 weights)."""

from __future__ import annotations

import torch

from . import _ops


class _SplineEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, u, kind):
        ctx.kind = kind
        ctx.shape = tuple(data.shape)
        u = u.detach().to(device=data.device, dtype=torch.float32).contiguous()
        ctx.save_for_backward(u)
        return _ops.spline_eval(data.detach().contiguous(), kind, u)

    @staticmethod
    def backward(ctx, grad_out):
        (u,) = ctx.saved_tensors
        grad = _ops.spline_eval_backward(ctx.shape, ctx.kind, u, grad_out.contiguous().to(torch.float32))
        return grad, None, None


class _CubicGrid3d(torch.nn.Module):
    grid_kind = 0

    def __init__(self, resolution=(2, 2, 2), n_channels: int = 1):
        super().__init__()
        if isinstance(resolution, int):
            resolution = (resolution,) * 3
        self._data = torch.nn.Parameter(torch.zeros((n_channels, *resolution), dtype=torch.float32))

    @property
    def data(self) -> torch.Tensor:
        return self._data

    @data.setter
    def data(self, value: torch.Tensor) -> None:
        self._data = torch.nn.Parameter(value)

    @property
    def resolution(self):
        return tuple(self._data.shape[1:])

    @property
    def n_channels(self) -> int:
        return self._data.shape[0]

    @classmethod
    def from_grid_data(cls, data: torch.Tensor):
        if data.ndim == 3:
            data = data[None]
        grid = cls(resolution=tuple(data.shape[1:]), n_channels=data.shape[0])
        grid._data = torch.nn.Parameter(data.detach().clone().to(torch.float32))
        return grid

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        if not self._data.is_cuda:
            raise RuntimeError("spline grids are evaluated by CUDA kernels only: move the grid with .to('cuda') first")
        return _SplineEval.apply(self._data, u, self.grid_kind)


class CubicCatmullRomGrid3d(_CubicGrid3d):
    grid_kind = 0


class CubicBSplineGrid3d(_CubicGrid3d):
    grid_kind = 1
