"""Drop-in for the reference's all-pairs correlation block and its tensor helpers.

Mirrors:
  * CorrBlock                       model/corr.py:12-60
  * bilinear_sampler, coords_grid, upflow8   model/model_utils.py:7-32
"""
from __future__ import annotations

import os

import torch

from . import autograd as ag
from . import ops


def _default_precision(D: int, H: int, W: int) -> str:
    """The reference forms the volume with fp32 torch.matmul (model/corr.py:58), where TF32 is off unless the
    user sets torch.backends.cuda.matmul.allow_tf32.  The drop-in follows the same switch: exact fp32 by default,
    the tcgen05 TF32 kernel when the user allowed TF32 (or asked for it with precision= / the environment
    variable EEMFLOW_B200_CORR_PRECISION = fp32 | tf32 | tf32_f16) and the shape is one the tensor-core path takes."""
    env = os.environ.get("EEMFLOW_B200_CORR_PRECISION")
    if env:
        return env
    if torch.backends.cuda.matmul.allow_tf32 and ops.tf32_supported(D, H, W):
        return "tf32"
    return "fp32"


class CorrBlock:
    """RAFT all-pairs correlation: volume + avg-pool pyramid at construction, window lookup on call.

    `corr_pyramid` keeps the reference's attribute contract: a list of `[B*H*W, 1, H_l, W_l]`
    float32 tensors.  `precision` is the one extra knob:
      "fp32"      exact fp32 FMA (the default, like the reference's fp32 matmul)
      "tf32"      tcgen05 TF32 tensor cores, f32 pyramid tensors
      "tf32_f16"  tcgen05 TF32 tensor cores, fp32 accumulation, the pyramid kept as the fp16 WORKING pyramid
                  (4x4-pixel tiles, all levels of a position in one row: half the bytes written by the GEMM and read
                  by every lookup).  `corr_pyramid` is then materialised lazily -- only if somebody reads the
                  attribute.  Inference only: under autograd this falls back to "tf32".
    """

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, precision=None):
        self.num_levels = num_levels
        self.radius = radius
        batch, dim, ht, wd = fmap1.shape
        if precision is None:
            precision = _default_precision(dim, ht, wd)
        if precision not in ("fp32", "tf32", "tf32_f16"):
            raise ValueError(f"eemflow_b200.CorrBlock: unknown precision {precision!r}")
        self._packed = None
        self._pyramid = None
        self._shape = (batch, ht, wd)
        if ag.needs_grad(fmap1, fmap2):
            if precision == "tf32_f16":
                precision = "tf32"
            self._pyramid = list(ag.CorrPyramidFn.apply(fmap1, fmap2, num_levels, precision))
        else:
            with torch.no_grad():
                if precision == "tf32_f16":
                    self._packed = ops.corr_pyramid_packed(fmap1, fmap2, num_levels)
                else:
                    self._pyramid = ops.corr_pyramid(fmap1, fmap2, num_levels, precision=precision)
        self.precision = precision

    @property
    def corr_pyramid(self):
        if self._pyramid is None:           # fp16 working pyramid: build the reference-shaped f32 tensors on demand
            with torch.no_grad():
                self._pyramid = ops.corr_pyramid_unpack(self._packed, *self._shape, self.num_levels)
        return self._pyramid

    @corr_pyramid.setter
    def corr_pyramid(self, value):
        self._pyramid, self._packed = list(value), None

    def __call__(self, coords):
        if torch.is_grad_enabled() and coords.requires_grad:     # never a silent zero gradient
            return self._lookup_differentiable_coords(coords)
        if self._packed is not None:
            with torch.no_grad():
                return ops.corr_lookup_packed(self._packed, coords, self.num_levels, self.radius)
        if ag.needs_grad(*self.corr_pyramid):
            return ag.CorrLookupFn.apply(coords, self.radius, *self.corr_pyramid)
        with torch.no_grad():
            return ops.corr_lookup(self.corr_pyramid, coords, self.radius)

    def _lookup_differentiable_coords(self, coords):
        """Coordinates that require grad: the reference's own composition (model/corr.py:29-50) over the differentiable
        bilinear sampler, so they receive the gradient F.grid_sample gives them there (and the pyramid its own).  Every
        shipped caller detaches the coordinates (model/eraft.py:141) and takes the fused lookup kernel instead."""
        r = self.radius
        coords = coords.permute(0, 2, 3, 1)
        batch, h1, w1, _ = coords.shape
        d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        delta = torch.stack(torch.meshgrid(d, d, indexing='ij'), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
        out_pyramid = []
        for i, corr in enumerate(self.corr_pyramid):
            if corr.numel() == 0:                           # a level pooled away entirely: zeros, like the fused lookup
                out_pyramid.append(coords.new_zeros(batch, h1, w1, (2 * r + 1) ** 2))
                continue
            coords_lvl = coords.reshape(batch * h1 * w1, 1, 1, 2) / 2 ** i + delta
            out_pyramid.append(ag.bilinear_sample(corr, coords_lvl).view(batch, h1, w1, -1))
        return torch.cat(out_pyramid, dim=-1).permute(0, 3, 1, 2).contiguous().float()

    @staticmethod
    def corr(fmap1, fmap2, precision=None):
        batch, dim, ht, wd = fmap1.shape
        if precision is None:
            precision = _default_precision(dim, ht, wd)
        if ag.needs_grad(fmap1, fmap2):
            (vol,) = ag.CorrPyramidFn.apply(fmap1, fmap2, 1, precision)
        else:
            with torch.no_grad():
                (vol,) = ops.corr_pyramid(fmap1, fmap2, 1, precision=precision)
        return vol.view(batch, ht, wd, 1, ht, wd)


def bilinear_sampler(img, coords, mode='bilinear', mask=False):
    """ Wrapper for grid_sample, uses pixel coordinates """
    if mode != 'bilinear':
        raise NotImplementedError("eemflow_b200.bilinear_sampler implements mode='bilinear' only")
    return ag.bilinear_sample(img, coords, mask=mask)     # differentiable w.r.t. img and coords, like F.grid_sample


def coords_grid(batch, ht, wd, device=None):
    """model/model_utils.py:24-27; `device` (extra, optional) builds the grid there directly, which keeps a model
    forward free of host->device copies and therefore capturable in a CUDA graph."""
    coords = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing='ij')
    coords = torch.stack(coords[::-1], dim=0).float()
    return coords[None].repeat(batch, 1, 1, 1)


def upflow8(flow, mode='bilinear'):
    if mode != 'bilinear':
        raise NotImplementedError("eemflow_b200.upflow8 implements mode='bilinear' only")
    new_size = (8 * flow.shape[2], 8 * flow.shape[3])
    return ag.bilinear_resize(flow, new_size, align_corners=True, scale0=8.0, scale1=8.0, scale_rest=8.0)
