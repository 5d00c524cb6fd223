"""Functional torch-tensor front end of the C ABI (one function per kernel family).

Every function takes CUDA tensors, allocates the output with torch (caller-owned memory, as the
ABI requires), launches on torch's current stream and returns without synchronising.  The
reference-shaped classes in event_utils.py / corr.py / correlation.py / warp.py are thin
wrappers over these.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Sequence

import torch

from . import _lib as L


def _dev(t: torch.Tensor) -> torch.device:
    return t.device


# ------------------------------------------------------------------------------------------ K1/K2
def voxelize(events: torch.Tensor, offsets: torch.Tensor, max_events: int, num_bins: int, height: int,
             width: int, *, normalize: bool = True, deterministic: bool = False,
             dropped: torch.Tensor | None = None, out: torch.Tensor | None = None,
             stats_out: torch.Tensor | None = None, timestamp_multiplier: float = 1.0) -> torch.Tensor:
    """Voxelize concatenated event windows.
    timestamp_multiplier: EventSequence's multiplier (loader/loader_utils.py:367-368) applied on the device.

    events  : CUDA float64 [N, 4] rows (ts, x, y, p)  -- the reference's EventSequence.features layout
    offsets : CUDA int64 [n_windows + 1]
    returns : CUDA float32 [n_windows, num_bins, height, width]
    Follows utils/transformers.py:56-122 per window.
    """
    assert events.dim() == 2 and events.shape[1] == 4, "events must be [N, 4]"
    assert num_bins > 0
    assert width > 0
    assert height > 0
    assert events.dtype == torch.float64, "Timestamps must be float64!"
    events = L.require_cuda(events, "events", torch.float64)
    offsets = L.require_cuda(offsets, "offsets", torch.int64)
    dev = _dev(events)
    n_windows = offsets.numel() - 1
    n_total = events.shape[0]
    if out is None:
        out = torch.empty((n_windows, num_bins, height, width), dtype=torch.float32, device=dev)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
        assert out.numel() == n_windows * num_bins * height * width
    lib = L.lib()
    mode = L.VOXEL_DETERMINISTIC if deterministic else L.VOXEL_ATOMIC
    with torch.cuda.device(dev):
        ws_bytes = lib.eem_voxelize_workspace_bytes(n_total, n_windows, num_bins, height, width, mode, int(normalize))
        ws = L.workspace.get(dev, ws_bytes, "voxel")
        L.check(lib.eem_voxelize_scaled(events.data_ptr(), float(timestamp_multiplier), offsets.data_ptr(), n_windows, n_total,
                                        int(max_events), num_bins, height, width, mode, int(normalize), out.data_ptr(),
                                        L.ptr(dropped), L.ptr(stats_out), L.ptr(ws), ws_bytes, L.stream_ptr(dev)))
    return out


def voxelize_soa(t: torch.Tensor, x: torch.Tensor, y: torch.Tensor, p: torch.Tensor, offsets: torch.Tensor,
                 max_events: int, num_bins: int, height: int, width: int, *, normalize: bool = True,
                 deterministic: bool = False, dropped: torch.Tensor | None = None,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """Voxelize packed event columns (13 B/event): t float64 [N] (values of features[:,0]) or int64 [N]
    raw nanoseconds (HREM .npz; the kernel applies *1e-9, *1e6 and the first-stamp subtraction of
    loader/loader_utils.py:34,367-397 in float64), x/y int16 [N], p int8 [N]."""
    assert t.dtype in (torch.float64, torch.int64), "t must be float64 or int64 (ns)"
    assert num_bins > 0 and width > 0 and height > 0
    t = L.require_cuda(t, "t", t.dtype)
    x = L.require_cuda(x, "x", torch.int16)
    y = L.require_cuda(y, "y", torch.int16)
    p = L.require_cuda(p, "p", torch.int8)
    offsets = L.require_cuda(offsets, "offsets", torch.int64)
    dev = t.device
    n_windows, n_total = offsets.numel() - 1, t.numel()
    assert x.numel() == n_total and y.numel() == n_total and p.numel() == n_total
    if out is None:
        out = torch.empty((n_windows, num_bins, height, width), dtype=torch.float32, device=dev)
    lib = L.lib()
    mode = L.VOXEL_DETERMINISTIC if deterministic else L.VOXEL_ATOMIC
    with torch.cuda.device(dev):
        ws_bytes = lib.eem_voxelize_workspace_bytes(n_total, n_windows, num_bins, height, width, mode, int(normalize))
        ws = L.workspace.get(dev, ws_bytes, "voxel")
        L.check(lib.eem_voxelize_soa(t.data_ptr(), int(t.dtype == torch.int64), x.data_ptr(), y.data_ptr(), p.data_ptr(),
                                     offsets.data_ptr(), n_windows, n_total, int(max_events), num_bins, height, width,
                                     mode, int(normalize), out.data_ptr(), L.ptr(dropped), None, L.ptr(ws), ws_bytes,
                                     L.stream_ptr(dev)))
    return out


def voxel_normalize_(grid: torch.Tensor, stats_out: torch.Tensor | None = None) -> torch.Tensor:
    """In-place non-zero mean/std normalisation of [n_windows, ...] grids (utils/transformers.py:114-122)."""
    grid = L.require_cuda(grid, "grid")
    dev = _dev(grid)
    n_windows = grid.shape[0]
    vox = grid.numel() // max(n_windows, 1)
    lib = L.lib()
    with torch.cuda.device(dev):
        ws_bytes = lib.eem_voxel_normalize_workspace_bytes(n_windows, vox)
        ws = L.workspace.get(dev, ws_bytes, "norm")
        L.check(lib.eem_voxel_normalize(grid.data_ptr(), n_windows, vox, L.ptr(stats_out), L.ptr(ws), ws_bytes,
                                        L.stream_ptr(dev)))
    return grid


# ------------------------------------------------------------------------------------------ K3/K4
def pyramid_level_shapes(h: int, w: int, num_levels: int) -> list[tuple[int, int]]:
    shapes = []
    for _ in range(num_levels):
        shapes.append((h, w))
        h, w = h // 2, w // 2
    return shapes


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, *, precision: str = "tf32",
                 out: Sequence[torch.Tensor] | None = None) -> list[torch.Tensor]:
    """All-pairs correlation pyramid (model/corr.py:13-27): list of [B*H*W, 1, H_l, W_l] float32."""
    fmap1 = L.require_cuda(fmap1, "fmap1")
    fmap2 = L.require_cuda(fmap2, "fmap2")
    assert fmap1.shape == fmap2.shape and fmap1.dim() == 4
    B, D, H, W = fmap1.shape
    dev = _dev(fmap1)
    shapes = pyramid_level_shapes(H, W, num_levels)
    if out is None:
        out = [torch.empty((B * H * W, 1, h, w), dtype=torch.float32, device=dev) for (h, w) in shapes]
    prec = {"fp32": L.CORR_FP32, "tf32": L.CORR_TF32}[precision]
    lib = L.lib()
    with torch.cuda.device(dev):
        ws_bytes = lib.eem_corr_pyramid_workspace_bytes(B, D, H, W, num_levels)
        ws = L.workspace.get(dev, ws_bytes, "corr")
        L.check(lib.eem_corr_pyramid(fmap1.data_ptr(), fmap2.data_ptr(), B, D, H, W, num_levels, L.ptr_array(out),
                                     prec, L.ptr(ws), ws_bytes, L.stream_ptr(dev)))
    return list(out)


def tf32_supported(D: int, H: int, W: int) -> bool:
    """Shapes the tcgen05 path takes (others use the fp32 kernel): see eem_corr_pyramid."""
    return (H * W) % 4 == 0 and D % 32 == 0 and D <= 256


def packed_row_elems(H: int, W: int, num_levels: int) -> tuple[int, list[int], list[int]]:
    """(elements per source position, level offsets, padded level sizes) of the fp16 working pyramid."""
    off = (C.c_int64 * num_levels)()
    ln = (C.c_int64 * num_levels)()
    row = C.c_int64(0)
    L.check(L.lib().eem_corr_packed_layout(H, W, num_levels, off, ln, C.byref(row)))
    return int(row.value), [int(v) for v in off], [int(v) for v in ln]


def corr_pyramid_packed(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4,
                        out: torch.Tensor | None = None) -> torch.Tensor:
    """All-pairs correlation pyramid as the fp16 WORKING pyramid: `[B*H*W, row]` float16, every level of a source
    position in one row, 4x4-pixel tiles (csrc/packed_layout.cuh).  TF32 tensor-core contraction with fp32
    accumulation, rounded once to fp16 on the way out.  Feed it to `corr_lookup_packed`; `corr_pyramid_unpack`
    recovers the reference's f32 `corr_pyramid` tensors (model/corr.py:13-27) when somebody wants to look at them."""
    fmap1 = L.require_cuda(fmap1, "fmap1")
    fmap2 = L.require_cuda(fmap2, "fmap2")
    assert fmap1.shape == fmap2.shape and fmap1.dim() == 4
    B, D, H, W = fmap1.shape
    dev = _dev(fmap1)
    row, _, _ = packed_row_elems(H, W, num_levels)
    if out is None:
        out = torch.empty((B * H * W, row), dtype=torch.float16, device=dev)
    else:
        assert out.is_cuda and out.dtype == torch.float16 and out.is_contiguous() and tuple(out.shape) == (B * H * W, row)
    lib = L.lib()
    with torch.cuda.device(dev):
        ws_bytes = lib.eem_corr_pyramid_packed_workspace_bytes(B, D, H, W, num_levels)
        ws = L.workspace.get(dev, ws_bytes, "corr_packed")
        L.check(lib.eem_corr_pyramid_packed(fmap1.data_ptr(), fmap2.data_ptr(), B, D, H, W, num_levels, out.data_ptr(),
                                            L.ptr(ws), ws_bytes, L.stream_ptr(dev)))
    return out


def corr_lookup_packed(packed: torch.Tensor, coords: torch.Tensor, num_levels: int, radius: int = 4,
                       out: torch.Tensor | None = None) -> torch.Tensor:
    """CorrBlock.__call__ (model/corr.py:29-50) on the fp16 working pyramid: coords [B,2,H,W] -> [B, L*(2r+1)^2, H, W] f32."""
    coords = L.require_cuda(coords, "coords")
    B, two, H, W = coords.shape
    assert two == 2
    assert packed.is_cuda and packed.dtype == torch.float16 and packed.is_contiguous()
    assert packed.shape[0] == B * H * W, "packed pyramid does not match coords"
    k = (2 * radius + 1) ** 2
    if out is None:
        out = torch.empty((B, num_levels * k, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        L.check(L.lib().eem_corr_lookup_packed(packed.data_ptr(), B, H, W, num_levels, radius, coords.data_ptr(),
                                               out.data_ptr(), L.stream_ptr(coords.device)))
    return out


def corr_pyramid_unpack(packed: torch.Tensor, B: int, H: int, W: int, num_levels: int) -> list[torch.Tensor]:
    """fp16 working pyramid -> the reference's list of [B*H*W, 1, H_l, W_l] float32 tensors."""
    assert packed.is_cuda and packed.dtype == torch.float16 and packed.is_contiguous()
    dev = packed.device
    out = [torch.empty((B * H * W, 1, h, w), dtype=torch.float32, device=dev) for (h, w) in pyramid_level_shapes(H, W, num_levels)]
    with torch.cuda.device(dev):
        L.check(L.lib().eem_corr_pyramid_unpack(packed.data_ptr(), B, H, W, num_levels, L.ptr_array(out), L.stream_ptr(dev)))
    return out


def avg_pool2x2(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(x, 2, stride=2) for [N, C, h, w] (model/corr.py:25-27)."""
    x = L.require_cuda(x, "x")
    n, c, h, w = x.shape
    out = torch.empty((n, c, h // 2, w // 2), dtype=torch.float32, device=x.device)
    if out.numel():
        with torch.cuda.device(x.device):
            L.check(L.lib().eem_avg_pool2x2(x.data_ptr(), n * c, h, w, out.data_ptr(), L.stream_ptr(x.device)))
    return out


# ------------------------------------------------------------------------------------------ K5
def corr_lookup(pyramid: Sequence[torch.Tensor], coords: torch.Tensor, radius: int = 4,
                out: torch.Tensor | None = None) -> torch.Tensor:
    """CorrBlock.__call__ (model/corr.py:29-50): coords [B,2,H,W] -> [B, L*(2r+1)^2, H, W]."""
    coords = L.require_cuda(coords, "coords")
    B, two, H, W = coords.shape
    assert two == 2
    num_levels = len(pyramid)
    for lvl in pyramid:
        if lvl.numel():
            assert lvl.is_cuda and lvl.dtype == torch.float32 and lvl.is_contiguous()
    assert pyramid[0].shape[0] == B * H * W, "pyramid does not match coords"
    k = (2 * radius + 1) ** 2
    if out is None:
        out = torch.empty((B, num_levels * k, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        L.check(L.lib().eem_corr_lookup(L.ptr_array(pyramid), B, H, W, num_levels, radius, coords.data_ptr(),
                                        out.data_ptr(), L.stream_ptr(coords.device)))
    return out


def bilinear_sample(img: torch.Tensor, coords: torch.Tensor, mask: bool = False):
    """bilinear_sampler (model/model_utils.py:7-21): img [N,C,H,W], coords [N,Ho,Wo,2] pixels -> [N,C,Ho,Wo]."""
    img = L.require_cuda(img, "img")
    coords = L.require_cuda(coords, "coords")
    N, Cc, H, W = img.shape
    n2, Ho, Wo, two = coords.shape
    assert n2 == N and two == 2
    out = torch.empty((N, Cc, Ho, Wo), dtype=torch.float32, device=img.device)
    m = torch.empty((N, Ho, Wo, 1), dtype=torch.float32, device=img.device) if mask else None
    with torch.cuda.device(img.device):
        L.check(L.lib().eem_bilinear_sample(img.data_ptr(), coords.data_ptr(), N, Cc, H, W, Ho, Wo, out.data_ptr(),
                                            L.ptr(m), L.stream_ptr(img.device)))
    return (out, m) if mask else out


def bilinear_sample_backward(img: torch.Tensor, coords: torch.Tensor, grad_out: torch.Tensor, need_img: bool = True,
                             need_coords: bool = True):
    """Gradients of bilinear_sample w.r.t. img [N,C,H,W] and coords [N,Ho,Wo,2] (None where not needed)."""
    img = L.require_cuda(img, "img")
    coords = L.require_cuda(coords, "coords")
    grad_out = L.require_cuda(grad_out, "grad_out")
    N, Cc, H, W = img.shape
    _, Ho, Wo, _ = coords.shape
    assert tuple(grad_out.shape) == (N, Cc, Ho, Wo)
    g_img = torch.empty_like(img) if need_img else None
    g_coords = torch.empty_like(coords) if need_coords else None
    if not (need_img or need_coords):
        return None, None
    with torch.cuda.device(img.device):
        L.check(L.lib().eem_bilinear_sample_backward(img.data_ptr(), coords.data_ptr(), grad_out.data_ptr(), N, Cc, H, W, Ho, Wo,
                                                     L.ptr(g_img), L.ptr(g_coords), L.stream_ptr(img.device)))
    return g_img, g_coords


# ------------------------------------------------------------------------------------------ K6
LOCAL_CORR_PRECISIONS = ("fp32", "tf32")
_local_corr_precision: str | None = None


def set_local_corr_precision(precision: str | None) -> None:
    """Arithmetic of the local 9x9 correlation in inference: "fp32" (FFMA, what the reference's sampler computes;
    the default) or "tf32" (tcgen05 banded GEMM, ~1e-3 relative).  None restores the default, which the
    environment variable EEMFLOW_B200_LOCAL_CORR_PRECISION can override."""
    global _local_corr_precision
    if precision is not None and precision not in LOCAL_CORR_PRECISIONS:
        raise ValueError(f"local-correlation precision must be one of {LOCAL_CORR_PRECISIONS}, got {precision!r}")
    _local_corr_precision = precision


def local_corr_precision() -> str:
    if _local_corr_precision is not None:
        return _local_corr_precision
    env = os.environ.get("EEMFLOW_B200_LOCAL_CORR_PRECISION")
    if env:
        if env not in LOCAL_CORR_PRECISIONS:
            raise ValueError(f"EEMFLOW_B200_LOCAL_CORR_PRECISION must be one of {LOCAL_CORR_PRECISIONS}, got {env!r}")
        return env
    return "fp32"


def local_corr_tf32_supported(B: int, Cc: int, H: int, W: int, max_disp: int = 4) -> bool:
    return bool(L.lib().eem_local_corr_tf32_supported(B, Cc, H, W, max_disp))


def local_corr(f1: torch.Tensor, f2: torch.Tensor, max_disp: int = 4, index: Sequence[int] | None = None,
               scale: float = 1.0, out: torch.Tensor | None = None, precision: str | None = None) -> torch.Tensor:
    """Local (2*md+1)^2 correlation, optional fused channel select and scale -> [B, n_out, H, W].

    precision None follows `local_corr_precision()`.  "tf32" runs the tcgen05 kernel on the shapes it takes
    (W % 4 == 0) and the exact FFMA kernel on the others (the coarsest 5x6 level of the MVSEC pyramid)."""
    f1 = L.require_cuda(f1, "f1")
    f2 = L.require_cuda(f2, "f2")
    assert f1.shape == f2.shape and f1.dim() == 4
    B, Cc, H, W = f1.shape
    nd = (2 * max_disp + 1) ** 2
    if index is None:
        n_out, idx = nd, None
    else:
        n_out = len(index)
        idx = (C.c_int * n_out)(*[int(v) for v in index])
    if out is None:
        out = torch.empty((B, n_out, H, W), dtype=torch.float32, device=f1.device)
    precision = precision or local_corr_precision()
    if precision not in LOCAL_CORR_PRECISIONS:
        raise ValueError(f"local-correlation precision must be one of {LOCAL_CORR_PRECISIONS}, got {precision!r}")
    use_tc = (precision == "tf32" and local_corr_tf32_supported(B, Cc, H, W, max_disp)
              and f1.data_ptr() % 16 == 0 and f2.data_ptr() % 16 == 0)
    fn = L.lib().eem_local_corr_tf32 if use_tc else L.lib().eem_local_corr
    with torch.cuda.device(f1.device):
        L.check(fn(f1.data_ptr(), f2.data_ptr(), B, Cc, H, W, max_disp, idx, n_out,
                   float(scale), out.data_ptr(), L.stream_ptr(f1.device)))
    return out


# ------------------------------------------------------------------------------------------ K7
def backwarp(x: torch.Tensor, flow: torch.Tensor, convention: int, mask_mode: int = L.MASK_NONE,
             return_mask: bool = False):
    x = L.require_cuda(x, "x")
    flow = L.require_cuda(flow, "flow")
    B, Cc, H, W = x.shape
    assert flow.shape == (B, 2, H, W), f"flow must be [B,2,H,W] matching x, got {tuple(flow.shape)}"
    out = torch.empty_like(x)
    mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device) if return_mask else None
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_backwarp(x.data_ptr(), flow.data_ptr(), B, Cc, H, W, convention, mask_mode,
                                     out.data_ptr(), L.ptr(mask), L.stream_ptr(x.device)))
    return (out, mask) if return_mask else out


def warp_blend(flow_init: torch.Tensor, inter_flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """torch_warp(flow_init, inter_flow) * (1 - mask) + flow_init * mask  (cdc_utils.py:173)."""
    flow_init = L.require_cuda(flow_init, "flow_init")
    inter_flow = L.require_cuda(inter_flow, "inter_flow")
    mask = L.require_cuda(mask, "mask")
    B, two, H, W = flow_init.shape
    assert two == 2 and inter_flow.shape == flow_init.shape and mask.shape == (B, 1, H, W)
    out = torch.empty_like(flow_init)
    with torch.cuda.device(flow_init.device):
        L.check(L.lib().eem_warp_blend(flow_init.data_ptr(), inter_flow.data_ptr(), mask.data_ptr(), B, H, W,
                                       out.data_ptr(), L.stream_ptr(flow_init.device)))
    return out


def upsample_flow_warp(coarse_flow: torch.Tensor, x: torch.Tensor, scale0: float, scale1: float):
    """(upsample2d_flow_as(coarse_flow, x, if_rate) without its in-place side effect, WarpingLayer_no_div(x, that flow)) in one launch."""
    coarse_flow = L.require_cuda(coarse_flow, "coarse_flow")
    x = L.require_cuda(x, "x")
    B, Cc, H, W = x.shape
    assert coarse_flow.shape[0] == B and coarse_flow.shape[1] == 2
    h, w = coarse_flow.shape[2:]
    flow = torch.empty((B, 2, H, W), dtype=torch.float32, device=x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_upsample_flow_warp(coarse_flow.data_ptr(), h, w, float(scale0), float(scale1), x.data_ptr(), B, Cc, H, W,
                                               flow.data_ptr(), out.data_ptr(), L.stream_ptr(x.device)))
    return flow, out


def blend_flow_warp(flow_init: torch.Tensor, inter_flow: torch.Tensor, mask: torch.Tensor, x: torch.Tensor):
    """(cdc blend(flow_init, inter_flow, mask), EEMFlow_cdc.warp(x, that flow)) in one launch."""
    flow_init = L.require_cuda(flow_init, "flow_init")
    inter_flow = L.require_cuda(inter_flow, "inter_flow")
    mask = L.require_cuda(mask, "mask")
    x = L.require_cuda(x, "x")
    B, Cc, H, W = x.shape
    assert tuple(flow_init.shape) == (B, 2, H, W) and inter_flow.shape == flow_init.shape and tuple(mask.shape) == (B, 1, H, W)
    flow = torch.empty_like(flow_init)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_blend_flow_warp(flow_init.data_ptr(), inter_flow.data_ptr(), mask.data_ptr(), x.data_ptr(), B, Cc, H, W,
                                            flow.data_ptr(), out.data_ptr(), L.stream_ptr(x.device)))
    return flow, out


# ------------------------------------------------------------------------------------------ K8/K9
def bilinear_resize(x: torch.Tensor, size: tuple[int, int], align_corners: bool, scale0: float = 1.0,
                    scale1: float = 1.0, scale_rest: float = 1.0, out: torch.Tensor | None = None) -> torch.Tensor:
    """`out` (optional) receives the result in place of a fresh tensor; it may live in a PEER GPU's memory (a slot of
    another rank's result buffer mapped into this process): the kernel's stores then travel over NVLink."""
    x = L.require_cuda(x, "x")
    B, Cc, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, Cc, H, W)
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_bilinear_resize(x.data_ptr(), B, Cc, h, w, out.data_ptr(), H, W, int(bool(align_corners)),
                                            float(scale0), float(scale1), float(scale_rest), L.stream_ptr(x.device)))
    return out


def scale_uv_(flow: torch.Tensor, scale0: float, scale1: float) -> torch.Tensor:
    assert flow.is_cuda and flow.dtype == torch.float32 and flow.is_contiguous()
    B, Cc, h, w = flow.shape
    with torch.cuda.device(flow.device):
        L.check(L.lib().eem_scale_uv_inplace(flow.data_ptr(), B, Cc, h, w, float(scale0), float(scale1),
                                             L.stream_ptr(flow.device)))
    return flow


def bilinear_resize_multi(xs: Sequence[torch.Tensor], size: tuple[int, int], align_corners: bool, scales: Sequence[tuple[float, float]],
                          scale_rest: float = 1.0, outs: Sequence[torch.Tensor | None] | None = None) -> list[torch.Tensor]:
    """bilinear_resize of several [B, C, h_k, w_k] maps (same B, C) to one size in ONE launch; scales[k] = (scale0, scale1)."""
    xs = [L.require_cuda(x, "x") for x in xs]
    n = len(xs)
    B, Cc = xs[0].shape[:2]
    H, W = int(size[0]), int(size[1])
    assert all(x.shape[0] == B and x.shape[1] == Cc for x in xs) and len(scales) == n
    res = []
    for k in range(n):
        o = None if outs is None else outs[k]
        if o is None:
            o = torch.empty((B, Cc, H, W), dtype=torch.float32, device=xs[0].device)
        else:
            assert o.is_cuda and o.dtype == torch.float32 and o.is_contiguous() and tuple(o.shape) == (B, Cc, H, W)
        res.append(o)
    hs = (C.c_int * n)(*[int(x.shape[2]) for x in xs])
    ws = (C.c_int * n)(*[int(x.shape[3]) for x in xs])
    s0 = (C.c_float * n)(*[float(s[0]) for s in scales])
    s1 = (C.c_float * n)(*[float(s[1]) for s in scales])
    with torch.cuda.device(xs[0].device):
        L.check(L.lib().eem_bilinear_resize_multi(L.ptr_array(xs), hs, ws, n, B, Cc, L.ptr_array(res), H, W, int(bool(align_corners)),
                                                  s0, s1, float(scale_rest), L.stream_ptr(xs[0].device)))
    return res


def scale_uv_multi_(flows: Sequence[torch.Tensor], scales: Sequence[tuple[float, float]]) -> None:
    """In-place per-channel scale of channels 0/1 of several [B, C, h_k, w_k] maps in ONE launch."""
    n = len(flows)
    B, Cc = flows[0].shape[:2]
    for f in flows:
        assert f.is_cuda and f.dtype == torch.float32 and f.is_contiguous() and f.shape[0] == B and f.shape[1] == Cc
    hs = (C.c_int * n)(*[int(f.shape[2]) for f in flows])
    ws = (C.c_int * n)(*[int(f.shape[3]) for f in flows])
    s0 = (C.c_float * n)(*[float(s[0]) for s in scales])
    s1 = (C.c_float * n)(*[float(s[1]) for s in scales])
    with torch.cuda.device(flows[0].device):
        L.check(L.lib().eem_scale_uv_inplace_multi(L.ptr_array(flows), hs, ws, n, B, Cc, s0, s1, L.stream_ptr(flows[0].device)))


def replicate_pad(x: torch.Tensor, pad: Sequence[int]) -> torch.Tensor:
    """F.pad(x, [left, right, top, bottom], mode='replicate') for [B,C,H,W]."""
    x = L.require_cuda(x, "x")
    B, Cc, H, W = x.shape
    left, right, top, bottom = (int(v) for v in pad)
    out = torch.empty((B, Cc, H + top + bottom, W + left + right), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_replicate_pad(x.data_ptr(), B, Cc, H, W, left, right, top, bottom, out.data_ptr(),
                                          L.stream_ptr(x.device)))
    return out


def inv_sqrt_dim(d: int) -> float:
    return 1.0 / math.sqrt(float(d))


# ------------------------------------------------------------------------------------------ backward (f1)
def local_corr_backward(f1, f2, grad_out, max_disp=4, index=None, scale=1.0, need_f1=True, need_f2=True):
    f1, f2, grad_out = L.require_cuda(f1, "f1"), L.require_cuda(f2, "f2"), L.require_cuda(grad_out, "grad_out")
    B, Cc, H, W = f1.shape
    n_out = grad_out.shape[1]
    idx = None if index is None else (C.c_int * n_out)(*[int(v) for v in index])
    g1 = torch.empty_like(f1) if need_f1 else None
    g2 = torch.empty_like(f2) if need_f2 else None
    with torch.cuda.device(f1.device):
        L.check(L.lib().eem_local_corr_backward(f1.data_ptr(), f2.data_ptr(), grad_out.data_ptr(), B, Cc, H, W, max_disp,
                                                idx, n_out, float(scale), L.ptr(g1), L.ptr(g2), L.stream_ptr(f1.device)))
    return g1, g2


def backwarp_backward(x, flow, grad_out, convention, mask_mode=L.MASK_NONE, need_x=True, need_flow=True):
    x, flow, grad_out = L.require_cuda(x, "x"), L.require_cuda(flow, "flow"), L.require_cuda(grad_out, "grad_out")
    B, Cc, H, W = x.shape
    gx = torch.empty_like(x) if need_x else None
    gf = torch.empty_like(flow) if need_flow else None
    with torch.cuda.device(x.device):
        L.check(L.lib().eem_backwarp_backward(x.data_ptr(), flow.data_ptr(), grad_out.data_ptr(), B, Cc, H, W, convention,
                                              mask_mode, L.ptr(gx), L.ptr(gf), L.stream_ptr(x.device)))
    return gx, gf


def bilinear_resize_backward(grad_out, in_size, align_corners, scale0=1.0, scale1=1.0, scale_rest=1.0):
    grad_out = L.require_cuda(grad_out, "grad_out")
    B, Cc, H, W = grad_out.shape
    h, w = int(in_size[0]), int(in_size[1])
    gin = torch.empty((B, Cc, h, w), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        L.check(L.lib().eem_bilinear_resize_backward(grad_out.data_ptr(), B, Cc, h, w, H, W, int(bool(align_corners)),
                                                     float(scale0), float(scale1), float(scale_rest), gin.data_ptr(),
                                                     L.stream_ptr(grad_out.device)))
    return gin


def corr_lookup_backward(grad_out: torch.Tensor, coords: torch.Tensor, radius: int, num_levels: int) -> list[torch.Tensor]:
    """d(levels) of CorrBlock.__call__ given d(out) [B, L*(2r+1)^2, H, W]: list of [B*H*W, 1, H_l, W_l]."""
    grad_out = L.require_cuda(grad_out, "grad_out")
    coords = L.require_cuda(coords, "coords")
    B, _, H, W = coords.shape
    dev = coords.device
    grads = [torch.empty((B * H * W, 1, h, w), dtype=torch.float32, device=dev) for (h, w) in pyramid_level_shapes(H, W, num_levels)]
    with torch.cuda.device(dev):
        L.check(L.lib().eem_corr_lookup_backward(grad_out.data_ptr(), coords.data_ptr(), B, H, W, num_levels, radius,
                                                 L.ptr_array(grads), L.stream_ptr(dev)))
    return grads


def avg_pool2x2_backward_(grad_fine: torch.Tensor, grad_coarse: torch.Tensor, accumulate: bool = True) -> torch.Tensor:
    """grad_fine [N,C,h,w] (+)= unpool(grad_coarse [N,C,h//2,w//2]) / 4, in place."""
    grad_fine = L.require_cuda(grad_fine, "grad_fine")
    n, c, h, w = grad_fine.shape
    assert tuple(grad_coarse.shape) == (n, c, h // 2, w // 2)
    if grad_coarse.numel():
        grad_coarse = L.require_cuda(grad_coarse, "grad_coarse")
    with torch.cuda.device(grad_fine.device):
        L.check(L.lib().eem_avg_pool2x2_backward(L.ptr(grad_coarse) if grad_coarse.numel() else None, n * c, h, w,
                                                 grad_fine.data_ptr(), int(accumulate), L.stream_ptr(grad_fine.device)))
    return grad_fine


def batched_gemm_tf32_supported(A: torch.Tensor, B: torch.Tensor, b_transposed: bool) -> bool:
    """True when `batched_gemm_(..., precision="tf32")` can run this product on the tcgen05 kernel (shape and
    alignment rules of eem_batched_gemm_tf32; contiguous [batch, rows, cols] operands assumed)."""
    batch, M, K = A.shape
    N = B.shape[1] if b_transposed else B.shape[2]
    if A.data_ptr() % 16 or B.data_ptr() % 16:
        return False
    return bool(L.lib().eem_batched_gemm_tf32_supported(batch, M, N, K, A.shape[2], B.shape[2], M * K,
                                                        B.shape[1] * B.shape[2], int(b_transposed)))


def batched_gemm_tf32_multi_(C_: torch.Tensor, As, Bs, *, b_transposed: bool, alpha: float = 1.0,
                             accumulate: bool = False) -> torch.Tensor:
    """C_[b] = alpha * sum_s As[s][b] @ (Bs[s][b].T if b_transposed else Bs[s][b]) (+ C_[b]) in ONE launch of the tcgen05
    kernel (one K loop over the segments); every segment must satisfy batched_gemm_tf32_supported."""
    As = [L.require_cuda(a, "A") for a in As]
    Bs = [L.require_cuda(b, "B") for b in Bs]
    n = len(As)
    assert n == len(Bs) and 1 <= n <= 6
    batch, M, N = C_.shape
    assert C_.is_cuda and C_.dtype == torch.float32 and C_.is_contiguous()
    Ks = []
    for a, b in zip(As, Bs):
        assert a.dim() == b.dim() == 3 and a.shape[0] == b.shape[0] == batch and a.shape[1] == M
        K = a.shape[2]
        assert tuple(b.shape[1:]) == ((N, K) if b_transposed else (K, N))
        if not batched_gemm_tf32_supported(a, b, b_transposed):
            raise NotImplementedError("eemflow_b200.batched_gemm_tf32_multi_: a segment is not addressable by the tcgen05 kernel")
        Ks.append(K)
    i64 = C.c_int64 * n
    with torch.cuda.device(C_.device):
        L.check(L.lib().eem_batched_gemm_tf32_multi(
            L.ptr_array(As), L.ptr_array(Bs), C_.data_ptr(), n, batch, M, N, (C.c_int * n)(*Ks),
            i64(*[a.shape[2] for a in As]), i64(*[b.shape[2] for b in Bs]), N,
            i64(*[a.shape[1] * a.shape[2] for a in As]), i64(*[b.shape[1] * b.shape[2] for b in Bs]), M * N,
            int(b_transposed), float(alpha), int(accumulate), L.stream_ptr(C_.device)))
    return C_


def batched_gemm_(C_: torch.Tensor, A: torch.Tensor, B: torch.Tensor, *, b_transposed: bool, alpha: float = 1.0,
                  accumulate: bool = False, precision: str = "fp32") -> torch.Tensor:
    """C_[b] = alpha * A[b] @ (B[b].T if b_transposed else B[b]) (+ C_[b]), all [batch, rows, cols] contiguous.

    precision "fp32": exact FFMA kernel; "tf32": the tcgen05 kernel (TF32 operands, fp32 accumulate) where its shape
    rules hold, the FFMA kernel otherwise (e.g. a 9x11 level whose row pitch is not a multiple of 16 bytes)."""
    if precision not in ("fp32", "tf32"):
        raise ValueError(f"eemflow_b200.batched_gemm_: unknown precision {precision!r}")
    A = L.require_cuda(A, "A")
    B = L.require_cuda(B, "B")
    assert C_.is_cuda and C_.dtype == torch.float32 and C_.is_contiguous() and A.dim() == B.dim() == C_.dim() == 3
    batch, M, K = A.shape
    N = B.shape[1] if b_transposed else B.shape[2]
    assert B.shape[0] == batch and (B.shape[2] if b_transposed else B.shape[1]) == K and tuple(C_.shape) == (batch, M, N)
    fn = L.lib().eem_batched_gemm_f32
    if precision == "tf32" and batched_gemm_tf32_supported(A, B, b_transposed):
        fn = L.lib().eem_batched_gemm_tf32
    with torch.cuda.device(A.device):
        L.check(fn(A.data_ptr(), B.data_ptr(), C_.data_ptr(), batch, M, N, K, A.shape[2], B.shape[2],
                   N, M * K, B.shape[1] * B.shape[2], M * N, int(b_transposed), float(alpha),
                   int(accumulate), L.stream_ptr(A.device)))
    return C_
