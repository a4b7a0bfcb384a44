"""Autograd glue for the differentiable entry points (the gradients are CUDA kernels)."""

from __future__ import annotations

import torch

from . import _ops
from ._lib import call, ptr, query, stream_ptr


class WarpTwoGrids(torch.autograd.Function):
    """``correct_motion_two_grids(grad=True)``: warped stack with a graph to the NEW grid's coefficients.

    Reference: correct_motion.py:188-317 (autograd through two ``grid_sample`` calls and the spline
    module).  Backward: d out / d shift per pixel -> transpose of the bicubic lattice lookup ->
    transpose of the spline evaluation (``tmc_warp_lattice_backward`` + ``tmc_spline_eval_backward``)."""

    @staticmethod
    def forward(ctx, new_data, movie, base_data, pixel_spacing, new_kind, base_kind):
        coeffs = new_data.detach().to(device=movie.device, dtype=torch.float32).contiguous()
        t = movie.shape[0]
        gh, gw = coeffs.shape[-2:]
        lattice = _ops.spline_lattice(coeffs, new_kind, t, 10 * gh, 10 * gw, coeffs2=base_data, kind2=base_kind)
        out = torch.empty_like(movie)
        _ops.warp_lattice(movie, lattice, pixel_spacing, out_stack=out)
        ctx.save_for_backward(movie, lattice)
        ctx.meta = (tuple(coeffs.shape), new_kind, float(pixel_spacing), new_data.device, new_data.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        movie, lattice = ctx.saved_tensors
        shape, kind, px, src_device, src_dtype = ctx.meta
        t, h, w = movie.shape
        _, _, lh, lw = lattice.shape
        dev = movie.device
        go = grad_out.detach().to(device=dev, dtype=torch.float32).contiguous()
        grad_lattice = torch.empty_like(lattice)
        ws = torch.empty((2 * query("tmc_warp_workspace_floats", t, w, lh),), dtype=torch.float32, device=dev)
        tyx = torch.empty((t, lh, lw, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            call("tmc_warp_lattice_backward", ptr(movie), t, h, w, ptr(lattice), lh, lw, px, ptr(go), ptr(grad_lattice), ptr(ws),
                 stream_ptr(dev))
            call("tmc_lattice_tyx", t, 0, t, lh, lw, ptr(tyx), stream_ptr(dev))
        grad_points = grad_lattice.permute(0, 2, 3, 1).contiguous()  # (t, lh, lw, 2): one row per lattice node
        grad = _ops.spline_eval_backward(shape, kind, tyx, grad_points)
        return grad.to(device=src_device, dtype=src_dtype), None, None, None, None, None
