"""GPU parity: local 9x9 correlation (csrc/local_corr.cu), backward warps, flow resize, blend and
replicate padding (csrc/warp.cu) against the real reference's golden vectors and the CPU oracle.
fp32 paths, gate <= 1e-5 abs (stated per assert).
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


# ------------------------------------------------------------------------------------ local corr
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_local_corr_golden(golden, E, name):
    g = golden("local_corr")
    f1, f2 = _t(g[f"{name}__f1"]).cuda(), _t(g[f"{name}__f2"]).cuda()
    ref = g[f"{name}__cv"]
    b, c, h, w = f1.shape
    out = E.Correlation(4)(f1, f2)
    assert tuple(out.shape) == (b, 81, h, w)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5
    raw = E.SpatialCorrelationSampler(1, 9, 1, 0, 1)(f1, f2)
    assert tuple(raw.shape) == (b, 9, 9, h, w)
    assert np.abs(raw.view(b, 81, h, w).cpu().numpy() / c - ref).max() <= 1e-5
    for idx in (g["index_cdc"], g["index_eemflow"]):
        sel = E.correlation_select(f1, f2, torch.as_tensor(idx))
        assert tuple(sel.shape) == (b, len(idx), h, w)
        assert np.abs(sel.cpu().numpy() - ref[:, idx]).max() <= 1e-5


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 5, 6), (2, 64, 40, 48), (1, 32, 80, 96), (1, 64, 96, 160), (1, 32, 192, 320), (1, 19, 33, 37),
                                     # whole-map-per-CTA kernel (<= 256 pixels): coarsest MVSEC / HREM levels, ragged, 1 x 1, the limit
                                     (32, 64, 10, 12), (2, 64, 12, 20), (3, 41, 7, 9), (2, 8, 1, 1), (1, 5, 16, 16), (1, 3, 2, 128)])
def test_local_corr_oracle_eemflow_shapes(E, B, C, H, W):
    """EEMFlow_cdc pyramid shapes at MVSEC (pad 320x384) and HREM (pad 768x1280), plus a ragged one."""
    gen = torch.Generator().manual_seed(C + H)
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)
    from eemflow_b200.correlation import EEMFLOW_CDC_INDEX
    ref = ref_ops.correlation(f1, f2, 4)
    out = E.Correlation(4)(f1.cuda(), f2.cuda()).cpu()
    assert (out - ref).abs().max().item() <= 1e-5
    sel = E.correlation_select(f1.cuda(), f2.cuda(), EEMFLOW_CDC_INDEX).cpu()
    assert torch.equal(sel, out[:, EEMFLOW_CDC_INDEX])


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 10, 12), (2, 64, 20, 24), (2, 64, 40, 48), (1, 32, 80, 96), (1, 64, 96, 160),
                                     (1, 19, 33, 36), (3, 40, 7, 100), (1, 8, 3, 4)])
def test_local_corr_tf32_tensor_core(E, B, C, H, W):
    """The tcgen05 banded-GEMM form of the local correlation (csrc/local_corr_tc.cu).  With inputs that are exactly
    representable in TF32 (low 13 mantissa bits cleared) the products are exact, so it must agree with the oracle
    like the FFMA kernel does (1e-5: accumulation order only) -- this pins every index of the band extraction,
    the zero padding and the ragged edges.  With full-precision inputs the stated tolerance is 2e-3 of the
    per-pixel feature energy |f1||f2|/C (TF32 rounds each factor to 10+1 bits: <= 2^-10 relative per product)."""
    from eemflow_b200 import ops
    from eemflow_b200.correlation import EEMFLOW_CDC_INDEX, EEMFLOW_INDEX
    gen = torch.Generator().manual_seed(7 * C + H)
    f1 = torch.randn(B, C, H, W, generator=gen)
    f2 = torch.randn(B, C, H, W, generator=gen)

    def tf32_exact(x):
        return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)

    a, b = tf32_exact(f1), tf32_exact(f2)
    ref = ref_ops.correlation(a, b, 4)
    assert ops.local_corr_tf32_supported(B, C, H, W)
    out = ops.local_corr(a.cuda(), b.cuda(), scale=1.0 / C, precision="tf32").cpu()
    assert (out - ref).abs().max().item() <= 1e-5
    for idx in (EEMFLOW_CDC_INDEX, EEMFLOW_INDEX, [80, 0, 40]):
        sel = ops.local_corr(a.cuda(), b.cuda(), index=idx, scale=1.0 / C, precision="tf32").cpu()
        assert torch.equal(sel, out[:, idx])
    # full-precision inputs: TF32 rounding only
    ref = ref_ops.correlation(f1, f2, 4)
    out = ops.local_corr(f1.cuda(), f2.cuda(), scale=1.0 / C, precision="tf32").cpu()
    energy = (f1.norm(dim=1, keepdim=True) * torch.nn.functional.max_pool2d(f2.norm(dim=1, keepdim=True), 9, 1, 4)) / C
    assert ((out - ref).abs() / energy.clamp_min(1e-6)).max().item() <= 2e-3
    exact = ops.local_corr(f1.cuda(), f2.cuda(), scale=1.0 / C, precision="fp32").cpu()
    assert (exact - ref).abs().max().item() <= 1e-5


def test_local_corr_precision_switch(E):
    """`set_local_corr_precision` routes Correlation / correlation_select; shapes the tensor-core kernel does not take
    (W % 4 != 0) run the exact kernel whatever the switch says."""
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(3)
    f1, f2 = torch.randn(2, 32, 12, 16, generator=gen).cuda(), torch.randn(2, 32, 12, 16, generator=gen).cuda()
    exact = E.Correlation(4)(f1, f2)
    try:
        E.set_local_corr_precision("tf32")
        assert E.local_corr_precision() == "tf32"
        fast = E.Correlation(4)(f1, f2)
        assert not torch.equal(fast, exact) and (fast - exact).abs().max().item() <= 5e-3
        g1, g2 = torch.randn(1, 16, 5, 6, generator=gen).cuda(), torch.randn(1, 16, 5, 6, generator=gen).cuda()
        assert not ops.local_corr_tf32_supported(1, 16, 5, 6)
        assert torch.equal(E.Correlation(4)(g1, g2), ops.local_corr(g1, g2, scale=1.0 / 16, precision="fp32"))
    finally:
        E.set_local_corr_precision(None)
    assert E.local_corr_precision() == "fp32"
    with pytest.raises(ValueError):
        E.set_local_corr_precision("bf16")


# ------------------------------------------------------------------------------------ warps
def _mask_knife_edge(x, flo):
    """The reference thresholds grid_sample(ones) at 1.0 (cdc_utils.py:77) / 0.9999 (tools.py:2251); for an
    interior sample that sum is 1 +- 1 ulp, so its own 0/1 mask is decided by rounding.  The kernels
    reproduce ATen's CPU rounding sequence (fma un-normalisation, weight products, add order), so the
    masks are expected to agree EXACTLY; this helper only marks those positions so a failure report
    can say whether a mismatch sits on such a knife edge."""
    _, raw = ref_ops.warping_layer_no_div(x, flo.clone(), return_raw_mask=True)
    return ((raw - 1.0).abs() <= 3e-7) | ((raw - 0.9999).abs() <= 3e-7)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_warp_golden(golden, E, name):
    g = golden("warp")
    x, flo = _t(g[f"{name}__x"]), _t(g[f"{name}__flo"])
    xc, fc = x.cuda(), flo.cuda()
    assert np.abs(E.warp(xc, fc).cpu().numpy() - g[f"{name}__warp_exact"]).max() <= 1e-5
    assert np.abs(E.tensor_tools.torch_warp(xc, fc).cpu().numpy() - g[f"{name}__torch_warp"]).max() <= 1e-5
    assert torch.equal(fc.cpu(), flo)                                          # flow is not modified
    edge = _mask_knife_edge(x, flo).numpy()
    o, m = E.tensor_tools.torch_warp_mask(xc, fc)
    assert tuple(m.shape) == tuple(x.shape)
    bad = m.cpu().numpy() != g[f"{name}__torch_warp_mask_mask"]
    assert not bad.any(), f"{bad.sum()} mask mismatches, {(bad & edge).sum()} of them on knife-edge sums"
    assert np.abs(o.cpu().numpy() - g[f"{name}__torch_warp_mask_out"]).max() <= 1e-5
    wl = E.WarpingLayer_no_div()(xc, fc).cpu().numpy()
    assert np.abs(wl - g[f"{name}__warping_layer"]).max() <= 1e-5          # includes the >= 1.0 mask, exactly


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 10, 12), (1, 32, 80, 96), (1, 64, 96, 160), (2, 2, 45, 80)])
def test_warp_oracle(E, B, C, H, W):
    gen = torch.Generator().manual_seed(H * W)
    x = torch.randn(B, C, H, W, generator=gen)
    flo = 2.0 * torch.randn(B, 2, H, W, generator=gen)
    xc, fc = x.cuda(), flo.cuda()
    assert (E.warp(xc, fc).cpu() - ref_ops.warp_exact(x, flo.clone())).abs().max().item() <= 1e-5
    assert (E.torch_warp(xc, fc).cpu() - ref_ops.torch_warp(x, flo.clone())).abs().max().item() <= 1e-5
    ref_wl, raw = ref_ops.warping_layer_no_div(x, flo.clone(), return_raw_mask=True)
    d = (E.WarpingLayer_no_div()(xc, fc).cpu() - ref_wl).abs()
    assert d.max().item() <= 1e-5, f"{(d > 1e-5).sum().item()} of {d.numel()} differ (mask threshold rounding?)"
    _, m = E.torch_warp_mask(xc, fc)
    assert torch.equal(m.cpu(), ref_ops.torch_warp_mask(x, flo.clone())[1])
    # identity flow under the exact convention returns the input up to the reference's own
    # normalise/un-normalise round trip (~W * eps in the coordinate, times the local gradient)
    zero = torch.zeros_like(fc)
    assert (E.warp(xc, zero) - xc).abs().max().item() <= 1e-4


def test_resize_blend_pad_golden(golden, E):
    g = golden("warp")
    tgt = torch.zeros(2, 1, 7, 9).cuda()
    a = _t(g["up_in"]).cuda()
    assert np.abs(E.upsample2d_flow_as(a, tgt, mode="bilinear", if_rate=False).cpu().numpy() - g["up_norate"]).max() <= 1e-5
    assert np.array_equal(a.cpu().numpy(), g["up_in"])
    b = _t(g["up_in"]).cuda()
    r = E.upsample2d_flow_as(b, tgt, mode="bilinear", if_rate=True)
    assert np.abs(r.cpu().numpy() - g["up_rate"]).max() <= 1e-5
    assert np.abs(b.cpu().numpy() - g["up_rate_input_after"]).max() <= 1e-6     # in-place side effect reproduced
    mesh = _t(g["mesh_in"]).cuda()
    assert np.abs(E.upsample_flow(mesh, (45, 80)).cpu().numpy() - g["mesh_up"]).max() <= 1e-5
    assert np.abs(E.upsample_flow(mesh, (5, 7)).cpu().numpy() - g["mesh_down"]).max() <= 1e-5
    bl = E.cdc_blend(_t(g["blend_init"]).cuda(), _t(g["blend_inter"]).cuda(), _t(g["blend_mask"]).cuda())
    assert np.abs(bl.cpu().numpy() - g["blend_out"]).max() <= 1e-5
    x = _t(g["pad_in"]).cuda()
    for mode, rate in (("chairs", 64), ("sintel", 32), ("chairs", 32)):
        p = E.InputPadder(x.shape, mode=mode, eval_pad_rate=rate)
        (y,) = p.pad(x)
        assert np.array_equal(y.cpu().numpy(), g[f"pad_{mode}_{rate}"])
        assert torch.equal(p.unpad(y), x)


@pytest.mark.parametrize("h,w,H,W", [(24, 40, 720, 1280), (16, 16, 720, 1280), (12, 20, 24, 40), (10, 12, 260, 346), (7, 9, 7, 9)])
def test_resize_oracle_meshflow_shapes(E, h, w, H, W):
    gen = torch.Generator().manual_seed(h * w)
    fl = torch.randn(2, 2, h, w, generator=gen)
    tgt = torch.zeros(2, 1, H, W)
    ref = ref_ops.upsample2d_flow_as(fl.clone(), tgt, if_rate=True)
    got = E.upsample2d_flow_as(fl.clone().cuda(), tgt.cuda(), if_rate=True).cpu()
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, W / w)            # values are scaled by W/w
    ref = ref_ops.upsample_flow(fl, (H, W))
    got = E.upsample_flow(fl.cuda(), (H, W)).cpu()
    assert (got - ref).abs().max().item() <= 1e-5


def test_upsample2d_flows_as_equals_the_list_of_calls():
    """upsample2d_flows_as (all predictions of EEMFlow_cdc in one launch, EEMFlow+.py:231-232) == the reference's list
    comprehension of upsample2d_flow_as calls: same outputs (<= 1e-5 against the CPU restatement, bit-identical to
    our single-map kernel) and the same in-place scaling of every input."""
    import eemflow_b200 as E
    gen = torch.Generator().manual_seed(12)
    shapes = [(5, 6), (10, 12), (20, 24), (40, 48), (80, 96)]
    for if_rate in (True, False):
        # (260, 346), (64, 96): upsampling or mixed targets -> row-walking kernel where every map is upsampled in y,
        # per-pixel kernel otherwise; (7, 9): a target smaller than most maps (source rows skip)
        for (H, W) in ((260, 346), (64, 96), (7, 9), (80, 8)):
            flows = [2.0 * torch.randn(3, 2, h, w, generator=gen) for (h, w) in shapes]
            tgt = torch.zeros(3, 1, H, W)
            ref_in = [f.clone() for f in flows]
            ref = [ref_ops.upsample2d_flow_as(f, tgt, if_rate=if_rate) for f in ref_in]
            a_in = [f.clone().cuda() for f in flows]
            b_in = [f.clone().cuda() for f in flows]
            multi = E.upsample2d_flows_as(a_in, tgt.cuda(), mode="bilinear", if_rate=if_rate)
            single = [E.upsample2d_flow_as(f, tgt.cuda(), mode="bilinear", if_rate=if_rate) for f in b_in]
            for k in range(len(flows)):
                assert torch.equal(multi[k], single[k]), k
                assert (multi[k].cpu() - ref[k]).abs().max().item() <= 1e-4 * max(1.0, ref[k].abs().max().item()), k
                assert torch.equal(a_in[k], b_in[k])
                assert (a_in[k].cpu() - ref_in[k]).abs().max().item() <= 1e-5 * max(1.0, ref_in[k].abs().max().item())
    # destination of the last map given by the caller
    out = torch.empty(3, 2, 64, 96, device="cuda")
    res = E.upsample2d_flows_as([f.clone().cuda() for f in flows], torch.zeros(3, 1, 64, 96, device="cuda"), if_rate=True, out_last=out)
    assert res[-1].data_ptr() == out.data_ptr()


def test_fused_level_chains_equal_the_separate_calls():
    """upsample_warp_no_div == upsample2d_flow_as + WarpingLayer_no_div and blend_warp == cdc_blend + warp
    (model/EEMFlow/cdc_utils.py:156-174, EEMFlow+.py:181): against the CPU restatement <= 1e-5 and against our own
    separate kernels (same arithmetic: <= 1e-6, masks identical); the deferred in-place scaling handed to
    upsample2d_flows_as(pre_scales=...) leaves outputs and inputs as the reference's call sequence does."""
    import eemflow_b200 as E
    gen = torch.Generator().manual_seed(21)
    for (B, C, h, w, H, W) in ((2, 32, 10, 12, 20, 24), (1, 64, 12, 20, 24, 40), (3, 5, 5, 6, 10, 12)):
        flow = 2.0 * torch.randn(B, 2, h, w, generator=gen)
        p1 = torch.randn(B, 32, H, W, generator=gen)
        p2 = torch.randn(B, 32, H, W, generator=gen)
        f2 = torch.randn(B, C, H, W, generator=gen)
        inter = 1.5 * torch.randn(B, 2, H, W, generator=gen)
        mask = torch.sigmoid(torch.randn(B, 1, H, W, generator=gen))
        # reference sequence (CPU restatement), including the in-place scaling of `flow`
        r_flow = flow.clone()
        r_up = ref_ops.upsample2d_flow_as(r_flow, p1, if_rate=True)
        r_w = ref_ops.warping_layer_no_div(p2, r_up)
        r_blend = ref_ops.cdc_blend(r_up, inter, mask)
        r_f2w = ref_ops.warp_exact(f2, r_blend.clone())
        # fused
        d_flow = flow.clone().cuda()
        up, wl, sc = E.upsample_warp_no_div(d_flow, p1.cuda(), p2.cuda())
        blend, f2w = E.blend_warp(up, inter.cuda(), mask.cuda(), f2.cuda())
        assert torch.equal(d_flow.cpu(), flow)                           # no side effect yet
        assert sc == (W / w, H / h)
        # separate kernels
        s_flow = flow.clone().cuda()
        s_up = E.upsample2d_flow_as(s_flow, p1.cuda(), mode="bilinear", if_rate=True)
        s_wl = E.WarpingLayer_no_div()(p2.cuda(), s_up)
        s_blend = E.cdc_blend(s_up, inter.cuda(), mask.cuda())
        s_f2w = E.warp(f2.cuda(), s_blend)
        # The >= 1.0 validity mask of WarpingLayer_no_div thresholds a weight sum that is 1 +- 1 ulp for interior samples,
        # so it is exact only for an IDENTICAL input flow (tests above); here the flow comes from our resize kernel, whose
        # last-ulp differences to ATen's interpolate flip isolated knife-edge pixels -- in the separate path just the same.
        for got, sep, ref, flips in ((up, s_up, r_up, 0.0), (wl, s_wl, r_w, 0.03), (blend, s_blend, r_blend, 0.0), (f2w, s_f2w, r_f2w, 0.0)):
            assert (got - sep).abs().max().item() <= 1e-6 * max(1.0, sep.abs().max().item())
            bad = ((got.cpu() - ref).abs() > 1e-4 * max(1.0, ref.abs().max().item())).any(dim=1)      # per pixel
            assert bad.float().mean().item() <= flips, (tuple(got.shape), bad.float().mean().item())
        assert torch.equal(wl == 0, s_wl == 0)
        # deferred scaling through the final upsampling: same prediction, same final state of the coarse flow
        tgt = torch.zeros(B, 1, 2 * H, 2 * W)
        r_final = ref_ops.upsample2d_flow_as(r_flow, tgt, if_rate=True)       # r_flow was scaled in place above, and is again here
        (final,) = E.upsample2d_flows_as([d_flow], tgt.cuda(), if_rate=True, pre_scales=[sc])
        assert (final.cpu() - r_final).abs().max().item() <= 1e-4 * max(1.0, r_final.abs().max().item())
        assert (d_flow.cpu() - r_flow).abs().max().item() <= 1e-5 * max(1.0, r_flow.abs().max().item())
