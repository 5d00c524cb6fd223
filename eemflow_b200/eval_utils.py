"""Loader- and evaluation-side helpers next to the hot path, on the device.

Mirrors (argument meaning and return values kept):
  * event_mask(event_sequence)                     loader/MVSEC.py:133-142   np.histogram2d(...) > 0
  * event_valid_from_volume(volume)                loader/HREM.py:238-239    np.sum(volume, axis=0)
  * flow_error(flow_gt, flow_pred, event_img, is_car=False, evaluation_type="sparse")
                                                   test_mvsec.py:291-346     Test.flow_error
  * motion_propagate(fflow, height, width, mesh_size=16, radius=3)
                                                   loader/HREM.py:41-101
plus batched forms (`flow_error_stats`, `motion_propagate_batch`) that keep everything on the GPU so an
evaluation loop reads back 40 bytes per sample instead of two full-resolution flow maps.
"""
from __future__ import annotations

import numpy
import torch

from . import _lib as L


def _cuda(t, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(numpy.ascontiguousarray(t))
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("eemflow_b200: no CUDA device is visible; this package has no CPU path")
        t = t.cuda()
    return t.to(dtype).contiguous()


def event_mask_batch(events: torch.Tensor, offsets: torch.Tensor, max_events: int, height: int, width: int) -> torch.Tensor:
    """events [N,4] float64 rows (device), offsets [n+1] int64 -> bool [n, 1, H, W]."""
    events = L.require_cuda(events, "events", torch.float64)
    offsets = L.require_cuda(offsets, "offsets", torch.int64)
    n = offsets.numel() - 1
    mask = torch.empty((n, 1, height, width), dtype=torch.uint8, device=events.device)
    with torch.cuda.device(events.device):
        L.check(L.lib().eem_event_mask(events.data_ptr(), offsets.data_ptr(), n, int(max_events), height, width,
                                       mask.data_ptr(), L.stream_ptr(events.device)))
    return mask.bool()


def event_mask(event_sequence) -> torch.Tensor:
    """return_dict['event_valid'] of the MVSEC validation loader: bool [1, H, W] on the GPU."""
    feats = event_sequence.get_sequence_only() if hasattr(event_sequence, "get_sequence_only") else event_sequence.features
    h, w = event_sequence.image_height, event_sequence.image_width
    ev = _cuda(feats, torch.float64)
    off = torch.tensor([0, ev.shape[0]], dtype=torch.int64, device=ev.device)
    return event_mask_batch(ev, off, ev.shape[0], h, w)[0]


def event_valid_from_volume(volume: torch.Tensor) -> torch.Tensor:
    """[nb, H, W] -> [1, H, W] (or [n, nb, H, W] -> [n, 1, H, W]): sum over bins, bin by bin in fp32."""
    v = _cuda(volume)
    single = v.dim() == 3
    if single:
        v = v[None]
    n, nb, h, w = v.shape
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        L.check(L.lib().eem_voxel_bin_sum(v.data_ptr(), n, nb, h, w, out.data_ptr(), L.stream_ptr(v.device)))
    return out[0] if single else out


def center_crop(x: torch.Tensor, size: int = 256) -> torch.Tensor:
    """The validation cropper of the MVSEC loader (loader/MVSEC.py:189-193: torchvision CenterCrop(256) on flow, voxel
    grids and the event mask): a view of the last two dimensions, no copy; the same rounding as torchvision
    (top = int(round((H - size) / 2)), left likewise)."""
    h, w = x.shape[-2:]
    top, left = int(round((h - size) / 2.0)), int(round((w - size) / 2.0))
    return x[..., top:top + size, left:left + size]


def flow_error_stats(flow_gt, flow_pred, event_img=None, max_row: int | None = None) -> torch.Tensor:
    """[B,2,H,W] x2 (+ [B,1,H,W] event image or None) -> float64 [B, 5] on the device:
    n_points, #(EE<1), #(EE<3 | EE<0.1|gt|), sum EE, sum |gt|."""
    gt, pred = _cuda(flow_gt), _cuda(flow_pred)
    assert gt.shape == pred.shape and gt.dim() == 4 and gt.shape[1] == 2
    B, _, H, W = gt.shape
    ev = None
    if event_img is not None:
        ev = _cuda(event_img).reshape(B, H, W)
    stats = torch.empty((B, 5), dtype=torch.float64, device=gt.device)
    with torch.cuda.device(gt.device):
        L.check(L.lib().eem_flow_error(gt.data_ptr(), pred.data_ptr(), L.ptr(ev), B, H, W, H if max_row is None else min(max_row, H),
                                       stats.data_ptr(), L.stream_ptr(gt.device)))
    return stats


def flow_error(flow_gt, flow_pred, event_img, is_car=False, evaluation_type="sparse"):
    """Test.flow_error: (AEE, percent_1_AEE, percent_3_AEE, n_points, AEE_sum, AEE_gt, AEE_gt_sum) of sample 0.

    `evaluation_type` stands for self.data_loader.dataset.evaluation_type.  The reference crops rows with
    `flow_gt.shape[1]` of the transposed [H, W, 2] array (i.e. W) unless is_car; kept.
    """
    gt, pred = flow_gt[:1], flow_pred[:1]
    H, W = gt.shape[-2:]
    max_row = 190 if is_car else W
    ev = None
    if evaluation_type == "sparse":
        ev = event_img[:1] if event_img.dim() == 4 else event_img.reshape(1, 1, H, W)
    n, c1, c3, s_ee, s_gt = flow_error_stats(gt, pred, ev, max_row)[0].tolist()
    n_points = int(n)
    percent_1 = float(c1 / float(n_points + 1e-5))
    percent_3 = float(c3) / float(n_points + 1e-5)
    if s_ee == 0:
        return 0, percent_1, percent_3, n_points, 0, 0, 0
    f32 = lambda x: torch.tensor(x, dtype=torch.float32)
    return f32(s_ee / n_points), percent_1, percent_3, n_points, f32(s_ee), f32(s_gt / n_points), f32(s_gt)


def motion_propagate_batch(fflow, mesh_size: int = 16, radius: int = 3) -> torch.Tensor:
    """[B, H, W, 2] float32 dense flow -> [B, 2, mesh, mesh] float32 mesh flow (x mesh, y mesh)."""
    f = _cuda(fflow)
    assert f.dim() == 4 and f.shape[-1] == 2
    B, H, W, _ = f.shape
    out = torch.empty((B, 2, mesh_size, mesh_size), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        L.check(L.lib().eem_motion_propagate(f.data_ptr(), B, H, W, mesh_size, radius, out.data_ptr(), L.stream_ptr(f.device)))
    return out


def motion_propagate(fflow, height, width, mesh_size=16, radius=3):
    """loader/HREM.py:41-101: [H, W, 2] -> (x_motion_mesh, y_motion_mesh) as float64 numpy [mesh, mesh]."""
    f = numpy.asarray(fflow)
    assert f.shape == (height, width, 2), "motion_propagate expects an [H, W, 2] flow field"
    m = motion_propagate_batch(torch.from_numpy(numpy.ascontiguousarray(f, dtype=numpy.float32))[None], mesh_size, radius)
    m = m[0].double().cpu().numpy()
    return m[0], m[1]
