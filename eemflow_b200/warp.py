"""Drop-ins for the reference's backward-warp, flow-resize and padding helpers.

Mirrors (names and argument meaning kept; every quirk of the sampling conventions kept, see
SURVEY.md section 0 trap 5):
  * warp(x, flo)                        EEMFlow_cdc.warp           model/EEMFlow/EEMFlow+.py:137-149
  * tensor_tools.torch_warp(x, flo)     utils_luo/tools.py:2262-2306
  * tensor_tools.torch_warp_mask(x,flo) utils_luo/tools.py:2217-2259
  * WarpingLayer_no_div()(x, flow)      model/EEMFlow/cdc_utils.py:50-78
  * upsample2d_flow_as(inputs, target_as, mode, if_rate)   model/EEMFlow/cdc_utils.py:80-103
  * upsample_flow(flow, orig_size)      EEMFlow.upsample_flow      model/EEMFlow/EEMFlow.py:118-120
  * cdc_blend(flow_init, inter_flow, inter_mask)           model/EEMFlow/cdc_utils.py:173
  * InputPadder                         utils/image_utils.py:126-145
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import autograd as ag
from . import ops


def warp(x, flo):
    """Bilinear sample of x at (px+u, py+v), zeros outside, exact pixel positions (align_corners=True)."""
    return ag.backwarp(x, flo, L.WARP_EXACT)


class tensor_tools:
    """The two classmethods of utils_luo/tools.py::tensor_tools that sit on the hot path."""

    @classmethod
    def torch_warp(cls, x, flo):
        """warp an image/tensor (im2) back to im1, according to the optical flow; x [B,C,H,W], flo [B,2,H,W].

        Keeps the reference's convention: (W-1) normalisation + grid_sample's default
        align_corners=False, i.e. the sample lands at px'*W/(W-1) - 0.5.
        """
        return ag.backwarp(x, flo, L.WARP_HALFPIX)

    @classmethod
    def torch_warp_mask(cls, x, flo):
        with torch.no_grad():
            out, mask = ops.backwarp(x, flo, L.WARP_HALFPIX, mask_mode=L.MASK_9999, return_mask=True)
        if ag.needs_grad(x, flo):  # differentiable output; the 0/1 mask is a constant, as in the reference
            out = ag.backwarp(x, flo, L.WARP_HALFPIX, L.MASK_9999)
        return out, mask.expand_as(out)


torch_warp = tensor_tools.torch_warp
torch_warp_mask = tensor_tools.torch_warp_mask


class WarpingLayer_no_div(nn.Module):
    def __init__(self):
        super(WarpingLayer_no_div, self).__init__()

    def forward(self, x, flow):
        return ag.backwarp(x, flow, L.WARP_HALFPIX, L.MASK_GE1)


def upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate=False, out=None):
    """Resize `inputs` to target_as's spatial size (align_corners=True); with if_rate scale u by w/w_ and
    v by h/h_ -- and, like the reference (cdc_utils.py:85-86), scale `inputs` itself IN PLACE.
    `out` (extra, optional): write the result there (inference only; may be peer-GPU memory, see ops.bilinear_resize)."""
    if mode != "bilinear":
        raise NotImplementedError("eemflow_b200.upsample2d_flow_as implements mode='bilinear' only")
    _, _, h, w = target_as.shape
    if if_rate:
        _, _, h_, w_ = inputs.shape
        u_scale, v_scale = (w / w_), (h / h_)
        res = ag.bilinear_resize(inputs, (h, w), align_corners=True, scale0=u_scale, scale1=v_scale, out=out)
        if inputs.is_contiguous() and inputs.dtype == torch.float32 and not ag.needs_grad(inputs):
            ops.scale_uv_(inputs, u_scale, v_scale)
        else:  # autograd-tracked or exotic view: keep the side effect with (tiny) torch in-place ops
            inputs[:, 0, :, :] *= u_scale
            inputs[:, 1, :, :] *= v_scale
        return res
    return ag.bilinear_resize(inputs, (h, w), align_corners=True, out=out)


def upsample2d_flows_as(inputs, target_as, mode="bilinear", if_rate=False, out_last=None, pre_scales=None):
    """`[upsample2d_flow_as(f, target_as, mode, if_rate) for f in inputs]` -- the five final predictions of EEMFlow_cdc
    (model/EEMFlow/EEMFlow+.py:231-232) -- in two launches instead of ten: one kernel resizes all maps, one applies the
    reference's in-place scaling of every input (cdc_utils.py:85-86).  Same results, same side effect.
    `out_last` (optional): destination of the LAST map's result (may be peer-GPU memory, see ops.bilinear_resize).
    `pre_scales` (optional): per map `(su, sv)` or None -- an in-place scaling of that input which an EARLIER
    upsample2d_flow_as(if_rate=True) call of the reference would already have applied (the per-level call of
    cdc_utils.py:156-160) and which the caller deferred (`upsample_warp_no_div`): it is folded into this call's resize
    and into its in-place scaling, so the inputs end up exactly as the reference leaves them."""
    if mode != "bilinear":
        raise NotImplementedError("eemflow_b200.upsample2d_flows_as implements mode='bilinear' only")
    inputs = list(inputs)
    pre = [(1.0, 1.0) if (pre_scales is None or pre_scales[k] is None) else tuple(pre_scales[k]) for k in range(len(inputs))]
    same = len({(tuple(f.shape[:2]), f.device, f.dtype) for f in inputs}) == 1
    if (not same or len(inputs) > 8 or ag.needs_grad(*inputs)
            or not all(f.is_cuda and f.is_contiguous() and f.dtype == torch.float32 for f in inputs)):
        for f, (su, sv) in zip(inputs, pre):           # apply the deferred scalings the way the reference would have
            if (su, sv) != (1.0, 1.0):
                f[:, 0, :, :] *= su
                f[:, 1, :, :] *= sv
        outs = [upsample2d_flow_as(f, target_as, mode, if_rate) for f in inputs[:-1]]
        return outs + [upsample2d_flow_as(inputs[-1], target_as, mode, if_rate, out=out_last)]
    _, _, h, w = target_as.shape
    rates = [((w / f.shape[3]), (h / f.shape[2])) if if_rate else (1.0, 1.0) for f in inputs]
    scales = [(r[0] * q[0], r[1] * q[1]) for r, q in zip(rates, pre)]
    with torch.no_grad():
        res = ops.bilinear_resize_multi(inputs, (h, w), True, scales, outs=[None] * (len(inputs) - 1) + [out_last])
        if any(sc != (1.0, 1.0) for sc in scales):
            ops.scale_uv_multi_(inputs, scales)
    return res


def upsample_warp_no_div(flow, feature_as, feature_to_warp):
    """`flow_up = upsample2d_flow_as(flow, feature_as, if_rate=True)` then `WarpingLayer_no_div()(feature_to_warp, flow_up)`
    (model/EEMFlow/cdc_utils.py:156-160) in ONE launch.  Returns (flow_up, warped, (su, sv)).  The reference's in-place
    scaling of `flow` by (su, sv) is NOT applied here: hand the pair to `upsample2d_flows_as(pre_scales=...)` (or scale
    `flow` yourself) if `flow` is used again, as EEMFlow_cdc does for its final predictions."""
    _, _, h, w = feature_as.shape
    _, _, h_, w_ = flow.shape
    su, sv = (w / w_), (h / h_)
    if ag.needs_grad(flow, feature_to_warp) or not (flow.is_cuda and feature_to_warp.is_cuda):
        flow_up = ag.bilinear_resize(flow, (h, w), align_corners=True, scale0=su, scale1=sv)
        return flow_up, ag.backwarp(feature_to_warp, flow_up, L.WARP_HALFPIX, L.MASK_GE1), (su, sv)
    with torch.no_grad():
        flow_up, warped = ops.upsample_flow_warp(flow, feature_to_warp, su, sv)
    return flow_up, warped, (su, sv)


def blend_warp(flow_init, inter_flow, inter_mask, x):
    """`flow_up = cdc_blend(flow_init, inter_flow, inter_mask)` (cdc_utils.py:173) then `warp(x, flow_up)`
    (EEMFlow+.py:137-149) in ONE launch.  Returns (flow_up, warped)."""
    if ag.needs_grad(flow_init, inter_flow, inter_mask, x):
        flow_up = cdc_blend(flow_init, inter_flow, inter_mask)
        return flow_up, warp(x, flow_up)
    with torch.no_grad():
        return ops.blend_flow_warp(flow_init, inter_flow, inter_mask, x)


def upsample_flow(flow, orig_size):
    """Meshflow -> dense flow: bilinear, align_corners=False, no magnitude rescale."""
    return ag.bilinear_resize(flow, tuple(orig_size), align_corners=False)


def cdc_blend(flow_init, inter_flow, inter_mask):
    """torch_warp(flow_init, inter_flow) * (1 - inter_mask) + flow_init * inter_mask, fused."""
    if ag.needs_grad(flow_init, inter_flow, inter_mask):
        # training: same expression as the reference, the warp itself runs (forward and backward) in the kernels
        return ag.backwarp(flow_init, inter_flow, L.WARP_HALFPIX) * (1 - inter_mask) + flow_init * inter_mask
    with torch.no_grad():
        return ops.warp_blend(flow_init, inter_flow, inter_mask)


class InputPadder:
    """ Pads images such that dimensions are divisible by eval_pad_rate (replicate padding) """

    def __init__(self, dims, mode='sintel', eval_pad_rate=32):
        self.eval_pad_rate = eval_pad_rate
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // eval_pad_rate) + 1) * eval_pad_rate - self.ht) % eval_pad_rate
        pad_wd = (((self.wd // eval_pad_rate) + 1) * eval_pad_rate - self.wd) % eval_pad_rate
        if mode == 'sintel':
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs):
        if ag.needs_grad(*inputs):      # the models pad event grids (no gradient); never return a silently detached result
            raise NotImplementedError("eemflow_b200.InputPadder.pad has no backward; pad tensors that do not require grad")
        with torch.no_grad():
            return [ops.replicate_pad(x, self._pad) for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        c = [self._pad[2], ht - self._pad[3], self._pad[0], wd - self._pad[1]]
        return x[..., c[0]:c[1], c[2]:c[3]]
