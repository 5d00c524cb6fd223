"""GPU parity: all-pairs correlation pyramid (csrc/corr_volume.cu) and window lookup
(csrc/corr_lookup.cu) against the real reference's golden vectors and the CPU oracle.

Gates: fp32 path <= 1e-5 abs on unit-variance volumes (stated per assert); TF32 tensor-core path
<= 1e-2*sigma per element (sigma = std of the reference volume) and <= 2e-3*sigma RMS.
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture(scope="module")
def E():
    import eemflow_b200
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    return eemflow_b200


@pytest.mark.parametrize("name", ["even", "odd", "wide", "degenerate"])
def test_golden_fp32_pyramid_and_lookup(golden, E, name):
    g = golden("corr")
    f1, f2, coords = _t(g[f"{name}__f1"]).cuda(), _t(g[f"{name}__f2"]).cuda(), _t(g[f"{name}__coords"]).cuda()
    L = int(g[f"{name}__levels"])
    blk = E.CorrBlock(f1, f2, num_levels=L, radius=4, precision="fp32")
    assert len(blk.corr_pyramid) == L
    for l, lvl in enumerate(blk.corr_pyramid):
        ref = g[f"{name}__pyr{l}"]
        assert tuple(lvl.shape) == ref.shape and lvl.dtype == torch.float32
        err = np.abs(lvl.cpu().numpy() - ref).max()
        assert err <= 1e-5, (name, l, err)
    vol = E.CorrBlock.corr(f1, f2, precision="fp32")
    assert tuple(vol.shape) == g[f"{name}__corr"].shape
    assert np.abs(vol.cpu().numpy() - g[f"{name}__corr"]).max() <= 1e-5
    out = blk(coords)
    ref = g[f"{name}__lookup"]
    assert tuple(out.shape) == ref.shape and out.is_contiguous()
    got = out.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref)), name       # degenerate 1-row level: NaN like the reference
    err = np.nanmax(np.abs(got - ref))
    assert err <= 2e-5, (name, err)                                 # 1e-5 from the volume + 1e-5 interpolation


@pytest.mark.parametrize("name", ["even", "wide"])
def test_golden_tf32_pyramid(golden, E, name):
    g = golden("corr")
    f1, f2 = _t(g[f"{name}__f1"]).cuda(), _t(g[f"{name}__f2"]).cuda()
    L = int(g[f"{name}__levels"])
    blk = E.CorrBlock(f1, f2, num_levels=L, radius=4, precision="tf32")
    for l, lvl in enumerate(blk.corr_pyramid):
        ref = g[f"{name}__pyr{l}"]
        sigma = ref.std()
        d = lvl.cpu().numpy() - ref
        assert np.abs(d).max() <= 1e-2 * sigma, (name, l, np.abs(d).max(), sigma)
        assert np.sqrt((d ** 2).mean()) <= 2e-3 * sigma, (name, l)


@pytest.mark.parametrize("B,D,H,W", [(2, 256, 36, 44), (1, 256, 32, 32), (1, 128, 23, 40), (3, 64, 17, 20)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_oracle_parity_mvsec_shapes(E, B, D, H, W, precision):
    """ERAFT feature-map shapes (260x346 -> 36x44, 256x256 crop -> 32x32) and ragged ones
    (P not a multiple of the 128-position tiles, odd level sizes)."""
    gen = torch.Generator().manual_seed(B * 1000 + H)
    f1 = torch.randn(B, D, H, W, generator=gen)
    f2 = torch.randn(B, D, H, W, generator=gen)
    coords = ref_ops.coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=gen)
    ref_pyr = ref_ops.corr_pyramid(f1, f2, 4)
    blk = E.CorrBlock(f1.cuda(), f2.cuda(), num_levels=4, radius=4, precision=precision)
    tol_abs = 1e-5 if precision == "fp32" else None
    for l, (lvl, ref) in enumerate(zip(blk.corr_pyramid, ref_pyr)):
        d = (lvl.cpu() - ref).abs()
        if tol_abs is not None:
            assert d.max().item() <= tol_abs, (l, d.max().item())
        else:
            sigma = ref.std().item()
            assert d.max().item() <= 1e-2 * sigma, (l, d.max().item(), sigma)
    # lookup on OUR pyramid vs the oracle lookup on the SAME pyramid isolates the gather kernel
    ours = blk(coords.cuda()).cpu()
    ref = ref_ops.corr_lookup([lvl.cpu() for lvl in blk.corr_pyramid], coords, 4)
    err = (ours - ref).abs().max().item()
    assert err <= 1e-5, err
    assert tuple(ours.shape) == (B, 324, H, W)


def test_bench_config_b32_tf32_full(E):
    """BASELINE configs[1] exactly as bench.py runs it: B = 32, D = 256, 36x44 (P = 1584), 4 levels, TF32.  Every
    element of every level against the CPU oracle (<= 1e-2 sigma max, <= 2e-3 sigma RMS), and the lookup of all
    32 samples against the oracle lookup on the same pyramid (<= 1e-5)."""
    B, D, H, W = 32, 256, 36, 44
    gen = torch.Generator().manual_seed(32)
    f1 = torch.randn(B, D, H, W, generator=gen)
    f2 = torch.randn(B, D, H, W, generator=gen)
    coords = ref_ops.coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=gen)
    ref_pyr = ref_ops.corr_pyramid(f1, f2, 4)
    blk = E.CorrBlock(f1.cuda(), f2.cuda(), num_levels=4, radius=4, precision="tf32")
    for l, (lvl, ref) in enumerate(zip(blk.corr_pyramid, ref_pyr)):
        d = lvl.cpu() - ref
        sigma = ref.std().item()
        assert d.abs().max().item() <= 1e-2 * sigma, (l, d.abs().max().item(), sigma)
        assert d.pow(2).mean().sqrt().item() <= 2e-3 * sigma, l
    ours = blk(coords.cuda()).cpu()
    ref = ref_ops.corr_lookup([lvl.cpu() for lvl in blk.corr_pyramid], coords, 4)
    assert (ours - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("B,D,H,W,L", [(2, 256, 36, 44, 4), (1, 256, 32, 32, 4), (1, 128, 23, 40, 4), (3, 64, 17, 20, 4),
                                       (1, 32, 12, 20, 2), (1, 32, 8, 2, 2), (32, 256, 36, 44, 4),
                                       (1, 32, 64, 64, 6), (2, 32, 32, 44, 1)])      # 6 levels: three lookup stages exceed the
def test_packed_fp16_pyramid_and_lookup(E, B, D, H, W, L):                           # shared memory -> one-batch kernel; 1 level
    """precision="tf32_f16": TF32 contraction stored as the fp16 working pyramid (4x4-pixel tiles, csrc/packed_layout.cuh).
    (1) the lazily materialised `corr_pyramid` against the CPU oracle, every element: <= 1e-2 sigma max and
    <= 2e-3 sigma RMS (TF32 inputs + one fp16 rounding of the fp32 accumulator); (2) the packed lookup against the
    oracle lookup on those same (fp16-valued) levels: <= 1e-5 -- the gather/interpolation is exact fp32 arithmetic on
    identical taps, including partial tiles, odd level sizes, out-of-map windows and the reference's NaN on a 1-wide
    level; (3) the BASELINE configs[1] size B = 32."""
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(B * 100 + H + L)
    f1 = torch.randn(B, D, H, W, generator=gen)
    f2 = torch.randn(B, D, H, W, generator=gen)
    coords = ref_ops.coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=gen)
    coords[0, :, 0, 0] = torch.tensor([-50.0, 3.0])           # a window entirely outside the map
    coords[0, :, -1, -1] = torch.tensor([W + 2.5, H - 0.25])   # straddling the right/bottom edge
    ref_pyr = ref_ops.corr_pyramid(f1, f2, L)
    blk = E.CorrBlock(f1.cuda(), f2.cuda(), num_levels=L, radius=4, precision="tf32_f16")
    assert blk._packed is not None and blk._packed.dtype == torch.float16
    row, offs, lens = ops.packed_row_elems(H, W, L)
    assert tuple(blk._packed.shape) == (B * H * W, row)
    ours_first = blk(coords.cuda()).cpu()                      # lookup BEFORE anybody touched corr_pyramid
    assert blk._pyramid is None                                # ... did not materialise the f32 tensors
    levels = [lvl.cpu() for lvl in blk.corr_pyramid]
    for l, (lvl, ref) in enumerate(zip(levels, ref_pyr)):
        assert tuple(lvl.shape) == tuple(ref.shape)
        if ref.numel() == 0:
            continue
        d = lvl - ref
        sigma = max(ref.std().item(), 1e-6)
        assert d.abs().max().item() <= 1e-2 * sigma, (l, d.abs().max().item(), sigma)
        assert d.pow(2).mean().sqrt().item() <= 2e-3 * sigma, l
    ref = ref_ops.corr_lookup(levels, coords, 4)
    assert tuple(ours_first.shape) == tuple(ref.shape)
    assert torch.equal(torch.isnan(ours_first), torch.isnan(ref))
    err = (ours_first - ref).nan_to_num(0.0).abs().max().item()
    assert err <= 1e-5, err
    # padding cells of the packed rows are exact zeros (partial tiles / level padding)
    pk = blk._packed.float().cpu().view(B * H * W, row)
    h, w = H, W
    for l in range(L):
        tx, ty = (w + 3) // 4, (h + 3) // 4
        blockv = pk[:, offs[l]:offs[l] + lens[l]]
        if h * w > 0:
            tiles = blockv[:, :tx * ty * 16].view(-1, ty, tx, 4, 4).permute(0, 1, 3, 2, 4).reshape(-1, ty * 4, tx * 4)
            assert torch.equal(tiles[:, :h, :w].reshape(B * H * W, 1, h, w), levels[l])
            assert tiles[:, h:, :].abs().max().item() == 0 if ty * 4 > h else True
            assert tiles[:, :, w:].abs().max().item() == 0 if tx * 4 > w else True
            assert blockv[:, tx * ty * 16:].abs().sum().item() == 0
        h, w = h // 2, w // 2


@pytest.mark.parametrize("B,H,W,L,r", [(12, 36, 44, 4, 4), (3, 17, 20, 4, 4), (2, 92, 160, 4, 4), (5, 24, 31, 3, 2), (1, 8, 8, 4, 1)])
def test_packed_lookup_warp_specialised_equals_one_batch_kernel(E, B, H, W, L, r):
    """The persistent warp-specialised lookup (geometry warps -> tile-copy warps -> interpolation warps, coupled by
    mbarriers over a ring of stages -- the default) runs the same per-element arithmetic as the one-batch kernel, so the two must agree BIT FOR BIT; (12, 36, 44) has more batches than 3 x SMs, so the stage ring wraps and
    the barrier phases flip; the other shapes have ragged last batches, three levels, small radii, degenerate levels."""
    import os
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(B * H + W)
    D = 32
    f1 = torch.randn(B, D, H, W, generator=gen).cuda()
    f2 = torch.randn(B, D, H, W, generator=gen).cuda()
    base = torch.stack(torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")[::-1], 0).float()
    coords = (base[None] + 4.0 * torch.randn(B, 2, H, W, generator=gen)).cuda()
    packed = ops.corr_pyramid_packed(f1, f2, L)
    outs = {}
    for cfg in ("0", "3x1x8x10", "2x1x4x5", "4x1x2x3"):   # stages x CTAs per SM x geometry warps x copy warps
        os.environ["EEM_LOOKUP_PACKED_WS"] = cfg
        try:
            outs[cfg] = ops.corr_lookup_packed(packed, coords, L, r).clone()
        finally:
            os.environ.pop("EEM_LOOKUP_PACKED_WS", None)
    outs["default"] = ops.corr_lookup_packed(packed, coords, L, r)
    for cfg, o in outs.items():
        assert torch.equal(torch.nan_to_num(o, nan=-7.0), torch.nan_to_num(outs["0"], nan=-7.0)), cfg


def test_packed_lookup_radius_variants(E):
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(17)
    f1 = torch.randn(1, 32, 12, 20, generator=gen)
    f2 = torch.randn(1, 32, 12, 20, generator=gen)
    coords = ref_ops.coords_grid(1, 12, 20) + 2.0 * torch.randn(1, 2, 12, 20, generator=gen)
    for L, r in ((1, 4), (2, 3), (3, 2), (2, 1), (4, 4)):
        blk = E.CorrBlock(f1.cuda(), f2.cuda(), num_levels=L, radius=r, precision="tf32_f16")
        ours = blk(coords.cuda()).cpu()
        ref = ref_ops.corr_lookup([lvl.cpu() for lvl in blk.corr_pyramid], coords, r)
        assert tuple(ours.shape) == tuple(ref.shape) == (1, L * (2 * r + 1) ** 2, 12, 20)
        assert torch.equal(torch.isnan(ours), torch.isnan(ref)), (L, r)
        assert (ours - ref).nan_to_num(0.0).abs().max().item() <= 1e-5, (L, r)


def test_lookup_radius_and_level_variants(E):
    gen = torch.Generator().manual_seed(7)
    f1 = torch.randn(1, 32, 12, 20, generator=gen)
    f2 = torch.randn(1, 32, 12, 20, generator=gen)
    coords = ref_ops.coords_grid(1, 12, 20) + 2.0 * torch.randn(1, 2, 12, 20, generator=gen)
    for L, r in ((1, 4), (2, 3), (3, 2), (2, 1)):
        blk = E.CorrBlock(f1.cuda(), f2.cuda(), num_levels=L, radius=r, precision="fp32")
        ours = blk(coords.cuda()).cpu()
        ref = ref_ops.corr_lookup(ref_ops.corr_pyramid(f1, f2, L), coords, r)
        assert tuple(ours.shape) == tuple(ref.shape) == (1, L * (2 * r + 1) ** 2, 12, 20)
        assert (ours - ref).abs().max().item() <= 2e-5, (L, r)


def test_full_size_hrem_volume_properties(E):
    """HREM size (P = 92*160 = 14720, D = 256; 867 MB level 0): too big for the CPU oracle in seconds,
    so check size-independent properties: (1) rows sampled at random against an fp64 dot product,
    (2) pooling consistency level_{l+1} == avg_pool2d(level_l), (3) bilinearity in fmap1."""
    B, D, H, W = 1, 256, 92, 160
    gen = torch.Generator().manual_seed(11)
    f1 = torch.randn(B, D, H, W, generator=gen).cuda()
    f2 = torch.randn(B, D, H, W, generator=gen).cuda()
    blk = E.CorrBlock(f1, f2, num_levels=4, radius=4, precision="tf32")
    P = H * W
    lv0 = blk.corr_pyramid[0].view(P, P)
    rows = torch.randint(0, P, (64,), generator=gen).cuda()
    # EVERY element of level 0 against an fp64 contraction of the same features (torch fp64 GEMM as the checker,
    # 1024 rows at a time): <= 1e-2 sigma max, <= 2e-3 sigma RMS; sigma == 1 for unit-variance features
    f1d, f2d = f1.view(D, P).double(), f2.view(D, P).double()
    worst, sq = 0.0, 0.0
    for r0 in range(0, P, 1024):
        ref_rows = (f1d[:, r0:r0 + 1024].t() @ f2d) / 16.0
        d = (lv0[r0:r0 + 1024].double() - ref_rows)
        worst = max(worst, d.abs().max().item())
        sq += d.pow(2).sum().item()
    assert worst <= 1e-2, worst
    assert (sq / (P * P)) ** 0.5 <= 2e-3
    del f1d, f2d, ref_rows, d
    for l in range(3):                      # pooling consistency over ALL rows, 2048 at a time
        for r0 in range(0, P, 2048):
            pooled = torch.nn.functional.avg_pool2d(blk.corr_pyramid[l][r0:r0 + 2048], 2, stride=2)
            dd = (pooled - blk.corr_pyramid[l + 1][r0:r0 + 2048]).abs().max().item()
            assert dd <= 1e-2, (l, r0, dd)
    blk2 = E.CorrBlock(2.0 * f1, f2, num_levels=1, radius=4, precision="tf32")
    assert torch.allclose(blk2.corr_pyramid[0].view(P, P)[rows], 2.0 * lv0[rows], atol=1e-6)   # exact scaling by 2
    coords = (ref_ops.coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=gen)).cuda()
    out = blk(coords)
    assert tuple(out.shape) == (1, 324, H, W) and torch.isfinite(out).all()
    # the lookup at 2048 positions against the oracle fed with those rows only
    pos = torch.randint(0, P, (2048,), generator=gen)
    sub = [lvl[pos.cuda()].cpu() for lvl in blk.corr_pyramid]
    c = coords.view(2, P)[:, pos.cuda()].cpu().view(1, 2, 1, 2048)
    ref = ref_ops.corr_lookup(sub, c, 4)                                 # [1, 324, 1, 2048]
    got = out.view(324, P)[:, pos.cuda()].cpu().view(1, 324, 1, 2048)
    assert (got - ref).abs().max().item() <= 1e-5


def test_bilinear_sampler_coords_grid_upflow8(golden, E):
    g = golden("corr")
    s, m = E.bilinear_sampler(_t(g["bs_img"]).cuda(), _t(g["bs_coords"]).cuda(), mask=True)
    assert np.abs(s.cpu().numpy() - g["bs_out"]).max() <= 1e-5
    assert np.array_equal(m.cpu().numpy(), g["bs_mask"])
    up = E.upflow8(_t(g["up8_in"]).cuda())
    assert np.abs(up.cpu().numpy() - g["up8_out"]).max() <= 1e-5
    assert torch.equal(E.coords_grid(2, 5, 7), ref_ops.coords_grid(2, 5, 7))


def test_avg_pool_kernel(E):
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(37, 1, 9, 13, generator=gen)
    ref = torch.nn.functional.avg_pool2d(x, 2, stride=2)
    assert (ops.avg_pool2x2(x.cuda()).cpu() - ref).abs().max().item() <= 1e-6


def test_tf32_unsupported_shape_is_loud(E):
    f = torch.randn(1, 16, 9, 13).cuda()
    with pytest.raises(NotImplementedError):
        E.CorrBlock(f, f, precision="tf32")
    blk = E.CorrBlock(f, f, num_levels=3)          # default picks fp32 for this shape
    assert blk.precision == "fp32"


def test_second_device_same_process(E):
    """The > 48 KiB dynamic shared-memory opt-in is per device: run the kernels that need it on cuda:1 after cuda:0
    in one process (nn.DataParallel in train_EEMFlow_HREM.py:117 does exactly that).  Skipped on a 1-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    gen = torch.Generator().manual_seed(3)
    f1 = torch.randn(1, 64, 16, 24, generator=gen)
    f2 = torch.randn(1, 64, 16, 24, generator=gen)
    coords = ref_ops.coords_grid(1, 16, 24) + 2.0 * torch.randn(1, 2, 16, 24, generator=gen)
    a = torch.randn(1, 32, 24, 40, generator=gen)
    b = torch.randn(1, 32, 24, 40, generator=gen)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        blk = E.CorrBlock(f1.to(dev), f2.to(dev), num_levels=3, radius=4, precision="tf32")
        outs.append((blk(coords.to(dev)).cpu(), E.Correlation(4)(a.to(dev), b.to(dev)).cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
