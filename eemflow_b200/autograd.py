"""torch.autograd.Function wrappers: forward and backward both run in the sm_100a kernels.

Used by the reference-shaped classes whenever an input requires grad, so the drop-in is valid inside
the reference's training loop (train_mvsec.py:251-258: loss.backward() through the model).  Covered
ops: local correlation (+ fused channel select), the three backward-warp variants, bilinear flow
resize, and the all-pairs CorrBlock (pyramid + lookup).
"""
from __future__ import annotations

import os

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


def _backward_precision(forward_precision: str) -> str:
    """Precision of the pyramid's backward GEMMs: that of the forward unless EEMFLOW_B200_CORR_BACKWARD says otherwise."""
    forced = os.environ.get("EEMFLOW_B200_CORR_BACKWARD", "").strip().lower()
    if forced:
        if forced not in ("fp32", "tf32"):
            raise ValueError(f"EEMFLOW_B200_CORR_BACKWARD must be fp32 or tf32, got {forced!r}")
        return forced
    return "tf32" if forward_precision in ("tf32", "tf32_f16") else "fp32"


def _tc_operand(t: torch.Tensor) -> torch.Tensor:
    """[batch, rows, cols] operand as the tcgen05 GEMM can address it: the column count (row pitch) a multiple of 4
    elements and a 16-byte aligned base; a zero-padded copy only where needed."""
    pad_c = (-t.shape[2]) % 4
    if pad_c:
        return torch.nn.functional.pad(t, (0, pad_c))
    return t if t.data_ptr() % 16 == 0 else t.clone()


class CorrPyramidFn(torch.autograd.Function):
    """CorrBlock.__init__ under autograd (model/corr.py:13-27, 52-60): fmaps -> pyramid levels.

    Forward is the one fused kernel (level_l = fmap1^T pool^l(fmap2) / sqrt(D)).  Backward uses the same
    linearity: d fmap1 = sum_l pool^l(fmap2) . dV_l^T, d pool^l(fmap2) = fmap1 . dV_l, folded back through
    the pooling by the avg-pool backward kernel.  The products are batched GEMMs in this library's own kernels (the
    1/sqrt(D) scale and the sum over the levels are fused into them): exact fp32 FFMA (eem_batched_gemm_f32) after an
    fp32 forward; after a TF32 forward the tcgen05 kernel (eem_batched_gemm_tf32: TF32 operands, fp32 accumulate, the
    precision torch gives the backward of a TF32 matmul): d fmap1 as ONE launch whose K loop runs over all levels,
    d pool^l(fmap2) one launch per level; a level whose row pitch TMA cannot address (9 x 11) goes through a zero-padded
    copy.  EEMFLOW_B200_CORR_BACKWARD=fp32|tf32 overrides the choice.
    """

    @staticmethod
    def forward(ctx, fmap1, fmap2, num_levels, precision):
        ctx.save_for_backward(fmap1, fmap2)
        ctx.num_levels = num_levels
        ctx.backward_precision = _backward_precision(precision)
        return tuple(ops.corr_pyramid(fmap1, fmap2, num_levels, precision=precision))

    @staticmethod
    def backward(ctx, *grad_levels):
        f1, f2 = ctx.saved_tensors
        B, D, H, W = f1.shape
        P = H * W
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        scale = ops.inv_sqrt_dim(D)
        # the tcgen05 kernel's rules: M = D in whole 32-column TMEM chunks, 16-byte row pitch of the feature maps
        tc = ctx.backward_precision == "tf32" and D % 32 == 0 and D <= 256 and P % 4 == 0
        f1m = f1.contiguous().reshape(B, D, P)
        d2_levels = []
        a_segs, b_segs = [], []                                                 # d fmap1: one K segment per level
        f2l = f2.contiguous()
        for l, g in enumerate(grad_levels):
            hl, wl = f2l.shape[-2:]
            Pl = hl * wl
            d2 = None
            if g is not None and Pl > 0:
                G = g.contiguous().reshape(B, P, Pl)
                f2m = f2l.reshape(B, D, Pl)
                if tc:
                    # TMA needs 16-byte row pitches: a level such as 9 x 11 goes through zero-padded copies (its dV is a
                    # few percent of the pyramid's bytes); the zero columns add nothing to the sums
                    G, f2m = _tc_operand(G), _tc_operand(f2m)
                if need1:            # d1[b,d,i] = s * sum_l sum_j f2l[b,d,j] * G_l[b,i,j]
                    a_segs.append(f2m)
                    b_segs.append(G)
                if need2:            # d2[b,d,j] = s * sum_i f1[b,d,i] * G[b,i,j]
                    d2 = torch.empty((B, D, G.shape[2]), dtype=torch.float32, device=f1.device)
                    ops.batched_gemm_(d2, _tc_operand(f1m) if tc else f1m, G, b_transposed=False, alpha=scale,
                                      precision="tf32" if tc else "fp32")
                    if d2.shape[2] != Pl:
                        d2 = d2[:, :, :Pl].contiguous()
                    d2 = d2.view(B, D, hl, wl)
            d2_levels.append((d2, (hl, wl)))
            if l + 1 < len(grad_levels):
                f2l = ops.avg_pool2x2(f2l)
        d1 = None
        if need1 and a_segs:
            d1 = torch.empty((B, D, P), dtype=torch.float32, device=f1.device)
            if tc:                   # all levels in one launch: one K loop, d fmap1 written once
                for k in range(0, len(a_segs), 6):
                    ops.batched_gemm_tf32_multi_(d1, a_segs[k:k + 6], b_segs[k:k + 6], b_transposed=True, alpha=scale,
                                                 accumulate=k > 0)
            else:
                for k, (a_, b_) in enumerate(zip(a_segs, b_segs)):
                    ops.batched_gemm_(d1, a_, b_, b_transposed=True, alpha=scale, accumulate=k > 0)
        g2 = None
        if need2:
            for d2, (hl, wl) in reversed(d2_levels):         # fold the coarse gradients down to level 0
                if g2 is not None:
                    if d2 is None:
                        d2 = torch.empty((B, D, hl, wl), dtype=torch.float32, device=f1.device)
                        ops.avg_pool2x2_backward_(d2, g2, accumulate=False)
                    else:
                        ops.avg_pool2x2_backward_(d2, g2, accumulate=True)
                g2 = d2
            if g2 is None:
                g2 = torch.zeros_like(f2)
        g1 = None
        if need1:
            g1 = torch.zeros_like(f1) if d1 is None else d1.view(B, D, H, W)
        return g1, g2, None, None


class CorrLookupFn(torch.autograd.Function):
    """CorrBlock.__call__ under autograd: gradient to the pyramid levels (the coordinates are detached by
    every caller, model/eraft.py:141, and get none here)."""

    @staticmethod
    def forward(ctx, coords, radius, *levels):
        ctx.save_for_backward(coords)
        ctx.cfg = (radius, len(levels))
        return ops.corr_lookup(levels, coords, radius)

    @staticmethod
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        radius, n = ctx.cfg
        return (None, None, *ops.corr_lookup_backward(grad_out.contiguous(), coords, radius, n))


class BilinearSampleFn(torch.autograd.Function):
    """bilinear_sampler (model/model_utils.py:7-21) under autograd: gradients to the image and to the pixel
    coordinates, as F.grid_sample gives them in the reference; the optional in-range mask is not differentiable."""

    @staticmethod
    def forward(ctx, img, coords, mask):
        ctx.save_for_backward(img, coords)
        res = ops.bilinear_sample(img, coords, mask=mask)
        if mask:
            ctx.mark_non_differentiable(res[1])
            return res
        return res

    @staticmethod
    def backward(ctx, grad_out, *unused):
        img, coords = ctx.saved_tensors
        g_img, g_coords = ops.bilinear_sample_backward(img, coords, grad_out.contiguous(), need_img=ctx.needs_input_grad[0],
                                                       need_coords=ctx.needs_input_grad[1])
        return g_img, g_coords, None


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


def bilinear_resize(x, size, align_corners, scale0=1.0, scale1=1.0, scale_rest=1.0, out=None):
    if needs_grad(x):
        if out is not None:
            raise NotImplementedError("eemflow_b200: `out=` is an inference-only option of the flow resize")
        return ResizeFn.apply(x, tuple(size), align_corners, scale0, scale1, scale_rest)
    with torch.no_grad():
        return ops.bilinear_resize(x, size, align_corners, scale0, scale1, scale_rest, out=out)


def bilinear_sample(img, coords, mask=False):
    if needs_grad(img, coords):
        return BilinearSampleFn.apply(img, coords, bool(mask))
    with torch.no_grad():
        return ops.bilinear_sample(img, coords, mask=mask)
