"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT part of the product path.

A restatement of the reference's (boomluo02/EEMFlow) hot-path arithmetic with the same ATen
operators the reference itself calls (index_add_, matmul, avg_pool2d, grid_sample, interpolate),
so it reproduces the reference's CPU results operation for operation.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module; nothing under `eemflow_b200/` does.

Parity pin: every function here is checked against outputs of the REAL reference code (imported
from /root/reference by `oracle/gen_golden.py`, vectors committed under `tests/golden/`) in
`tests/test_oracle_golden.py`.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so those generated vectors are the pin.

Each function cites the reference lines it follows (paths relative to the reference tree).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# voxelization -- utils/transformers.py:36-124 (== utils_luo/event_utils.py:163-253,
#                                                  loader/loader_utils.py:447-537)
# ----------------------------------------------------------------------------------------------
def voxel_votes(features: np.ndarray, num_bins: int, height: int, width: int):
    """Per-event vote addresses and weights (utils/transformers.py:66-110), before accumulation.

    Returns dict of numpy arrays: idx_left/idx_right (int64 flat index, -1 where the reference's
    valid mask drops the vote), val_left/val_right (float32), tis (int64 floor of the normalised time).
    """
    events = torch.from_numpy(features.astype('float'))            # :46, :58
    last_stamp = events[-1, 0]                                       # :66
    first_stamp = events[0, 0]                                       # :67
    deltaT = last_stamp - first_stamp                                # :72
    if deltaT == 0:                                                  # :74-75
        deltaT = 1.0
    ts = (num_bins - 1) * (events[:, 0] - first_stamp) / deltaT      # :77
    xs = events[:, 1].long()                                         # :79
    ys = events[:, 2].long()                                         # :80
    pols = events[:, 3].float()                                      # :81
    pols[pols == 0] = -1                                             # :82
    tis = torch.floor(ts)                                            # :84
    tis_long = tis.long()                                            # :85
    dts = ts - tis                                                   # :86
    vals_left = pols * (1.0 - dts.float())                           # :87
    vals_right = pols * dts.float()                                  # :88
    valid_l = (tis < num_bins) & (tis >= 0)                          # :90-91
    valid_r = ((tis + 1) < num_bins) & (tis >= 0)                    # :104-105
    idx_l = xs + ys * width + tis_long * width * height              # :99-100
    idx_r = xs + ys * width + (tis_long + 1) * width * height        # :108-109
    idx_l = torch.where(valid_l, idx_l, torch.full_like(idx_l, -1))
    idx_r = torch.where(valid_r, idx_r, torch.full_like(idx_r, -1))
    return {
        "idx_left": idx_l.numpy(), "idx_right": idx_r.numpy(),
        "val_left": vals_left.numpy(), "val_right": vals_right.numpy(),
        "tis": tis_long.numpy(),
    }


def voxelize(features: np.ndarray, num_bins: int, height: int, width: int, normalize: bool = True) -> torch.Tensor:
    """EventSequenceToVoxelGrid_Pytorch.__call__ on the CPU (utils/transformers.py:36-124)."""
    assert features.shape[1] == 4                                    # :51
    assert num_bins > 0 and width > 0 and height > 0                 # :52-54
    v = voxel_votes(features, num_bins, height, width)
    grid = torch.zeros(num_bins, height, width, dtype=torch.float32).flatten()   # :63
    il, ir = torch.from_numpy(v["idx_left"]), torch.from_numpy(v["idx_right"])
    vl, vr = torch.from_numpy(v["val_left"]), torch.from_numpy(v["val_right"])
    ml, mr = il >= 0, ir >= 0
    grid.index_add_(dim=0, index=il[ml], source=vl[ml])              # :98-102  (all left votes, event order)
    grid.index_add_(dim=0, index=ir[mr], source=vr[mr])              # :107-110 (then all right votes)
    grid = grid.view(num_bins, height, width)                        # :112
    if normalize:                                                    # :114-122
        mask = torch.nonzero(grid, as_tuple=True)
        if mask[0].size()[0] > 0:
            mean = grid[mask].mean()
            std = grid[mask].std()
            if std > 0:
                grid[mask] = (grid[mask] - mean) / std
            else:
                grid[mask] = grid[mask] - mean
    return grid


# ----------------------------------------------------------------------------------------------
# all-pairs correlation -- model/corr.py:12-60, model/model_utils.py:7-32
# ----------------------------------------------------------------------------------------------
def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """CorrBlock.corr (model/corr.py:52-60) -> [B, h, w, 1, h, w]."""
    batch, dim, ht, wd = fmap1.shape
    f1 = fmap1.view(batch, dim, ht * wd)
    f2 = fmap2.view(batch, dim, ht * wd)
    corr = torch.matmul(f1.transpose(1, 2), f2)
    corr = corr.view(batch, ht, wd, 1, ht, wd)
    return corr / torch.sqrt(torch.tensor(dim).float())


def corr_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4) -> list[torch.Tensor]:
    """CorrBlock.__init__ (model/corr.py:13-27) -> list of [B*h*w, 1, h_l, w_l]."""
    corr = corr_volume(fmap1, fmap2)
    batch, h1, w1, dim, h2, w2 = corr.shape
    corr = corr.reshape(batch * h1 * w1, dim, h2, w2)
    pyramid = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyramid.append(corr)
    return pyramid


def bilinear_sampler(img: torch.Tensor, coords: torch.Tensor, mask: bool = False):
    """model/model_utils.py:7-21."""
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    grid = torch.cat([xgrid, ygrid], dim=-1)
    img = F.grid_sample(img, grid, align_corners=True)
    if mask:
        m = (xgrid > -1) & (ygrid > -1) & (xgrid < 1) & (ygrid < 1)
        return img, m.float()
    return img


def corr_lookup(pyramid: list[torch.Tensor], coords: torch.Tensor, radius: int = 4) -> torch.Tensor:
    """CorrBlock.__call__ (model/corr.py:29-50): coords [B,2,h,w] -> [B, L*(2r+1)^2, h, w]."""
    r = radius
    coords = coords.permute(0, 2, 3, 1)
    batch, h1, w1, _ = coords.shape
    out_pyramid = []
    for i, corr in enumerate(pyramid):
        dx = torch.linspace(-r, r, 2 * r + 1)
        dy = torch.linspace(-r, r, 2 * r + 1)
        delta = torch.stack(torch.meshgrid(dy, dx, indexing='ij'), axis=-1)
        centroid_lvl = coords.reshape(batch * h1 * w1, 1, 1, 2) / 2 ** i
        delta_lvl = delta.view(1, 2 * r + 1, 2 * r + 1, 2)
        coords_lvl = centroid_lvl + delta_lvl
        s = bilinear_sampler(corr, coords_lvl)
        out_pyramid.append(s.view(batch, h1, w1, -1))
    out = torch.cat(out_pyramid, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


def coords_grid(batch: int, ht: int, wd: int) -> torch.Tensor:
    """model/model_utils.py:24-27."""
    coords = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing='ij')
    coords = torch.stack(coords[::-1], dim=0).float()
    return coords[None].repeat(batch, 1, 1, 1)


def upflow8(flow: torch.Tensor) -> torch.Tensor:
    """model/model_utils.py:30-32."""
    new_size = (8 * flow.shape[2], 8 * flow.shape[3])
    return 8 * F.interpolate(flow, size=new_size, mode='bilinear', align_corners=True)


# ----------------------------------------------------------------------------------------------
# local 9x9 correlation -- EEMFlow.py:14-23 around spatial_correlation_sampler (not vendored,
# pinned spatial-correlation-sampler==0.4.0 in requirements.txt:131).  Restated after the in-tree
# equivalent model/IRRPWC/pwc_modules.py:42-63 (compute_cost_volume: zero pad, dy outer / dx
# inner, mean over channels == sum / C) and the dead CUDA kernel
# model/IRRPWC/correlation_package/correlation_cuda_kernel.cu:82-109.
# ----------------------------------------------------------------------------------------------
def local_corr_sampler(f1: torch.Tensor, f2: torch.Tensor, max_disp: int = 4) -> torch.Tensor:
    """SpatialCorrelationSampler(1, 2*md+1, 1, 0, 1)(f1, f2) -> [B, 2md+1, 2md+1, H, W] (channel sum, no /C)."""
    b, c, h, w = f1.shape
    n = 2 * max_disp + 1
    f2p = F.pad(f2, (max_disp, max_disp, max_disp, max_disp), "constant", 0)
    planes = []
    for i in range(n):          # dy
        for j in range(n):      # dx
            planes.append(torch.sum(f1 * f2p[:, :, i:(h + i), j:(w + j)], dim=1, keepdim=True))
    return torch.cat(planes, dim=1).view(b, n, n, h, w)


def correlation(f1: torch.Tensor, f2: torch.Tensor, max_disp: int = 4, index=None) -> torch.Tensor:
    """Correlation.forward (EEMFlow.py:21-23) [+ torch.index_select(.., dim=1, index) of EEMFlow.py:160]."""
    b, c, h, w = f1.shape
    out = local_corr_sampler(f1, f2, max_disp).view(b, -1, h, w) / c
    if index is not None:
        out = torch.index_select(out, dim=1, index=torch.as_tensor(index).long())
    return out


# ----------------------------------------------------------------------------------------------
# warps -- EEMFlow+.py:137-149, utils_luo/tools.py:2217-2306, cdc_utils.py:50-78
# ----------------------------------------------------------------------------------------------
def _vgrid(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.size()
    xx = torch.arange(0, W).view(1, -1).repeat(H, 1)
    yy = torch.arange(0, H).view(-1, 1).repeat(1, W)
    xx = xx.view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = yy.view(1, 1, H, W).repeat(B, 1, 1, 1)
    grid = torch.cat((xx, yy), 1).float()
    vgrid = grid + flo
    vgrid[:, 0, :, :] = 2.0 * vgrid[:, 0, :, :] / max(W - 1, 1) - 1.0
    vgrid[:, 1, :, :] = 2.0 * vgrid[:, 1, :, :] / max(H - 1, 1) - 1.0
    return vgrid.permute(0, 2, 3, 1)


def warp_exact(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    """EEMFlow_cdc.warp (EEMFlow+.py:137-149): align_corners=True."""
    return F.grid_sample(x, _vgrid(x, flo), mode='bilinear', align_corners=True)


def torch_warp(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    """tensor_tools.torch_warp (utils_luo/tools.py:2262-2306): grid_sample default align_corners=False."""
    return F.grid_sample(x, _vgrid(x, flo), padding_mode='zeros', align_corners=False)  # the reference relies on the default


def torch_warp_mask(x: torch.Tensor, flo: torch.Tensor):
    """tensor_tools.torch_warp_mask (utils_luo/tools.py:2217-2259)."""
    vgrid = _vgrid(x, flo)
    output = F.grid_sample(x, vgrid, padding_mode='zeros', align_corners=False)
    mask = F.grid_sample(torch.ones(x.size()), vgrid, padding_mode='zeros', align_corners=False)
    mask[mask < 0.9999] = 0
    mask[mask > 0] = 1
    return output * mask, mask


def warping_layer_no_div(x: torch.Tensor, flow: torch.Tensor, return_raw_mask: bool = False):
    """WarpingLayer_no_div.forward (cdc_utils.py:55-78)."""
    vgrid = _vgrid(x, flow)
    x_warp = F.grid_sample(x, vgrid, padding_mode='zeros', align_corners=False)
    raw = F.grid_sample(torch.ones(x.size()), vgrid, align_corners=False)
    mask = (raw >= 1.0).float()
    if return_raw_mask:
        return x_warp * mask, raw
    return x_warp * mask


def upsample2d_flow_as(inputs: torch.Tensor, target_as: torch.Tensor, if_rate: bool = False) -> torch.Tensor:
    """cdc_utils.py:80-103, including the in-place scaling of `inputs` (:85-86)."""
    _, _, h, w = target_as.shape
    res = F.interpolate(inputs, [h, w], mode="bilinear", align_corners=True)
    if if_rate:
        _, _, h_, w_ = inputs.shape
        inputs[:, 0, :, :] *= (w / w_)
        inputs[:, 1, :, :] *= (h / h_)
        u, v = res.chunk(2, dim=1)
        u = u * (w / w_)
        v = v * (h / h_)
        res = torch.cat([u, v], dim=1)
    return res


def upsample_flow(flow: torch.Tensor, size) -> torch.Tensor:
    """EEMFlow.upsample_flow (EEMFlow.py:118-120) == HREM GT upsample (loader/HREM.py:264-268)."""
    return F.interpolate(flow, size=tuple(size), mode='bilinear', align_corners=False)


def cdc_blend(flow_init: torch.Tensor, inter_flow: torch.Tensor, inter_mask: torch.Tensor) -> torch.Tensor:
    """cdc_model.forward blend (cdc_utils.py:173)."""
    return torch_warp(flow_init, inter_flow) * (1 - inter_mask) + flow_init * inter_mask


def input_pad(x: torch.Tensor, dims, mode: str = 'chairs', eval_pad_rate: int = 64):
    """InputPadder (utils/image_utils.py:126-145) -> (padded, pad list)."""
    ht, wd = dims[-2:]
    pad_ht = (((ht // eval_pad_rate) + 1) * eval_pad_rate - ht) % eval_pad_rate
    pad_wd = (((wd // eval_pad_rate) + 1) * eval_pad_rate - wd) % eval_pad_rate
    if mode == 'sintel':
        pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
    else:
        pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]
    return F.pad(x, pad, mode='replicate'), pad


# ------------------------------------------------------------------------------------------------
# loader / evaluation helpers
# ------------------------------------------------------------------------------------------------
def event_mask(features: np.ndarray, height: int, width: int) -> np.ndarray:
    """loader/MVSEC.py:133-142: pixels that received at least one event (bool [H, W])."""
    hist, _, _ = np.histogram2d(x=features[:, 1], y=features[:, 2], bins=(width, height), range=[[0, width], [0, height]])
    return hist.transpose() > 0


def voxel_bin_sum(volume: np.ndarray) -> np.ndarray:
    """loader/HREM.py:238-239: event_valid = np.sum(event_volume_old, axis=0)."""
    return np.sum(volume, axis=0)


def flow_error(flow_gt: torch.Tensor, flow_pred: torch.Tensor, event_img: torch.Tensor, is_car: bool = False,
               evaluation_type: str = "sparse"):
    """test_mvsec.py:291-346 (Test.flow_error), evaluation_type = self.data_loader.dataset.evaluation_type."""
    gt = flow_gt[0].numpy().transpose(1, 2, 0)
    pred = flow_pred[0].numpy().transpose(1, 2, 0)
    max_row = 190 if is_car else gt.shape[1]          # sic: the reference takes shape[1] of the [H, W, 2] array
    gt, pred = gt[:max_row, :], pred[:max_row, :]
    flow_mask = np.logical_and(np.logical_and(~np.isinf(gt[:, :, 0]), ~np.isinf(gt[:, :, 1])), np.linalg.norm(gt, axis=2) > 0)
    if evaluation_type == "sparse":
        event_mask_ = np.squeeze(event_img.numpy())[:max_row, :] > 0
        total = np.squeeze(np.logical_and(event_mask_, flow_mask))
    else:
        total = flow_mask
    gt_m, pred_m = gt[total, :], pred[total, :]
    EE = np.linalg.norm(gt_m - pred_m, axis=-1)
    EE_gt = np.linalg.norm(gt_m, axis=-1)
    n_points = EE.shape[0]
    percent_1 = float((EE < 1.).sum() / float(EE.shape[0] + 1e-5))
    percent_3 = float(((EE < 3.) | (EE < 0.1 * EE_gt)).sum()) / float(EE.shape[0] + 1e-5)
    EE, EE_gt = torch.from_numpy(EE), torch.from_numpy(EE_gt)
    if torch.sum(EE) == 0:
        return 0, percent_1, percent_3, n_points, 0, 0, 0
    return torch.mean(EE), percent_1, percent_3, n_points, torch.sum(EE), torch.mean(EE_gt), torch.sum(EE_gt)


def motion_propagate(fflow: np.ndarray, height: int, width: int, mesh_size: int = 16, radius: int = 3):
    """loader/HREM.py:30-101: [H, W, 2] dense flow -> two [mesh, mesh] float64 meshes (x, y)."""
    from scipy.signal import medfilt2d
    u, v = fflow[..., 0], fflow[..., 1]
    mesh_cols, mesh_rows = width // mesh_size, height // mesh_size
    clamp = lambda p, hi: min(max(p, 0), hi - 1)
    xm = np.zeros((mesh_size, mesh_size), dtype=float)
    ym = np.zeros((mesh_size, mesh_size), dtype=float)
    for i in range(mesh_size):
        for j in range(mesh_size):
            us, vs = [], []
            for r in range(radius):
                ox, oy = r * mesh_rows // 2, r * mesh_cols // 2
                for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                    pi, pj = clamp(mesh_rows * i + si * ox, height), clamp(mesh_cols * j + sj * oy, width)
                    us.append(u[pi, pj])
                    vs.append(v[pi, pj])
            if us:
                xm[i, j] = sorted(us)[len(us) // 2]
                ym[i, j] = sorted(vs)[len(vs) // 2]
    pad = 2
    xp, yp = np.pad(xm, pad, mode="edge"), np.pad(ym, pad, mode="edge")      # cv2.BORDER_REPLICATE
    xp, yp = medfilt2d(xp, [5, 5]), medfilt2d(yp, [5, 5])
    return xp[pad:pad + mesh_size, pad:pad + mesh_size], yp[pad:pad + mesh_size, pad:pad + mesh_size]
