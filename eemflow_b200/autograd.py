"""torch.autograd.Function wrappers: forward and backward both run in the sm_100a kernels.

Used by the reference-shaped classes whenever an input requires grad, so the drop-in is valid inside
the reference's training loop (train_mvsec.py:251-258: loss.backward() through the model).  Covered
ops: local correlation (+ fused channel select), the three backward-warp variants, bilinear flow
resize.  The all-pairs CorrBlock has no backward yet (its callers, the RAFT-family baselines, are
outside the EEMFlow training path); it raises if a gradient is requested.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops


class LocalCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, max_disp, index, scale):
        ctx.save_for_backward(f1, f2)
        ctx.cfg = (max_disp, None if index is None else tuple(int(v) for v in index), float(scale))
        return ops.local_corr(f1, f2, max_disp=max_disp, index=index, scale=scale)

    @staticmethod
    def backward(ctx, grad_out):
        f1, f2 = ctx.saved_tensors
        max_disp, index, scale = ctx.cfg
        g1, g2 = ops.local_corr_backward(f1, f2, grad_out.contiguous(), max_disp, index, scale,
                                         need_f1=ctx.needs_input_grad[0], need_f2=ctx.needs_input_grad[1])
        return g1, g2, None, None, None


class BackwarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow, convention, mask_mode):
        ctx.save_for_backward(x, flow)
        ctx.cfg = (convention, mask_mode)
        return ops.backwarp(x, flow, convention, mask_mode)

    @staticmethod
    def backward(ctx, grad_out):
        x, flow = ctx.saved_tensors
        convention, mask_mode = ctx.cfg
        gx, gf = ops.backwarp_backward(x, flow, grad_out.contiguous(), convention, mask_mode,
                                       need_x=ctx.needs_input_grad[0], need_flow=ctx.needs_input_grad[1])
        return gx, gf, None, None


class ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, align_corners, scale0, scale1, scale_rest):
        ctx.cfg = (tuple(x.shape[-2:]), bool(align_corners), float(scale0), float(scale1), float(scale_rest))
        return ops.bilinear_resize(x, size, align_corners, scale0, scale1, scale_rest)

    @staticmethod
    def backward(ctx, grad_out):
        in_size, align, s0, s1, sr = ctx.cfg
        return ops.bilinear_resize_backward(grad_out.contiguous(), in_size, align, s0, s1, sr), None, None, None, None, None


def needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def local_corr(f1, f2, max_disp=4, index=None, scale=1.0):
    if needs_grad(f1, f2):
        return LocalCorrFn.apply(f1, f2, max_disp, index, scale)
    with torch.no_grad():
        return ops.local_corr(f1, f2, max_disp=max_disp, index=index, scale=scale)


def backwarp(x, flow, convention, mask_mode=L.MASK_NONE):
    if needs_grad(x, flow):
        return BackwarpFn.apply(x, flow, convention, mask_mode)
    with torch.no_grad():
        return ops.backwarp(x, flow, convention, mask_mode)


def bilinear_resize(x, size, align_corners, scale0=1.0, scale1=1.0, scale_rest=1.0):
    if needs_grad(x):
        return ResizeFn.apply(x, tuple(size), align_corners, scale0, scale1, scale_rest)
    with torch.no_grad():
        return ops.bilinear_resize(x, size, align_corners, scale0, scale1, scale_rest)
