"""GPU parity of the backward kernels (csrc/backward.cu) and the autograd wrappers against torch autograd
through the CPU oracle (the reference's own differentiable ATen calls).  Gate: <= 2e-4 relative to the
largest gradient magnitude (fp32, differently ordered sums; the warp gradients multiply values ~N(0,1)).
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import eemflow_b200
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    return eemflow_b200


def close(a, b, tol=2e-4):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-6)
    return (a - b).abs().max().item() <= tol * scale, (a - b).abs().max().item() / scale


@pytest.mark.parametrize("B,C,H,W,select", [(2, 8, 7, 9, False), (1, 16, 12, 20, True), (2, 64, 20, 24, True), (1, 19, 33, 37, False)])
def test_local_corr_backward(E, B, C, H, W, select):
    from eemflow_b200.correlation import EEMFLOW_CDC_INDEX
    gen = torch.Generator().manual_seed(C * H)
    f1 = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    f2 = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    idx = EEMFLOW_CDC_INDEX if select else None
    ref = ref_ops.correlation(f1, f2, 4, index=idx)
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    a = f1.detach().cuda().requires_grad_(True)
    b = f2.detach().cuda().requires_grad_(True)
    out = E.correlation_select(a, b, idx) if select else E.Correlation(4)(a, b)
    assert out.requires_grad
    assert close(out, ref, 1e-5)[0]
    out.backward(g.cuda())
    ok1, e1 = close(a.grad, f1.grad)
    ok2, e2 = close(b.grad, f2.grad)
    assert ok1 and ok2, (e1, e2)
    # only one input needs a gradient
    a2 = f1.detach().cuda().requires_grad_(True)
    out2 = E.Correlation(4)(a2, f2.detach().cuda()) if not select else E.correlation_select(a2, f2.detach().cuda(), idx)
    out2.backward(g.cuda())
    assert close(a2.grad, f1.grad)[0]


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 7, 9), (1, 32, 20, 24), (2, 2, 45, 80)])
@pytest.mark.parametrize("variant", ["exact", "halfpix", "no_div"])
def test_backwarp_backward(E, B, C, H, W, variant):
    gen = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    flo = (2.0 * torch.randn(B, 2, H, W, generator=gen)).requires_grad_(True)
    ref_fn = {"exact": ref_ops.warp_exact, "halfpix": ref_ops.torch_warp, "no_div": ref_ops.warping_layer_no_div}[variant]
    ref = ref_fn(x, flo)
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    xc = x.detach().cuda().requires_grad_(True)
    fc = flo.detach().cuda().requires_grad_(True)
    fn = {"exact": E.warp, "halfpix": E.torch_warp, "no_div": E.WarpingLayer_no_div()}[variant]
    out = fn(xc, fc)
    assert out.requires_grad and close(out, ref, 1e-5)[0]
    out.backward(g.cuda())
    okx, ex = close(xc.grad, x.grad)
    okf, ef = close(fc.grad, flo.grad)
    assert okx and okf, (variant, ex, ef)


@pytest.mark.parametrize("h,w,H,W,align,rate", [(3, 5, 7, 9, True, True), (12, 20, 24, 40, True, True), (10, 12, 260, 346, True, True),
                                                 (16, 16, 45, 80, False, False), (24, 40, 5, 7, False, False), (7, 9, 7, 9, True, False)])
def test_resize_backward(E, h, w, H, W, align, rate):
    gen = torch.Generator().manual_seed(h * W)
    fl = torch.randn(2, 2, h, w, generator=gen, requires_grad=True)
    tgt = torch.zeros(2, 1, H, W)
    if align:
        ref = ref_ops.upsample2d_flow_as(fl.clone(), tgt, if_rate=rate)
    else:
        ref = ref_ops.upsample_flow(fl, (H, W))
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    fc = fl.detach().cuda().requires_grad_(True)
    if align:
        out = E.upsample2d_flow_as(fc.clone(), tgt.cuda(), if_rate=rate)
    else:
        out = E.upsample_flow(fc, (H, W))
    assert out.requires_grad
    out.backward(g.cuda())
    ok, e = close(fc.grad, fl.grad)
    assert ok, e


def test_cdc_blend_and_chain_backward(E):
    """A small differentiable chain as in cdc_model.forward + EEMFlow_cdc.forward:
    upsample (rate) -> WarpingLayer_no_div -> blend -> warp -> correlation."""
    from eemflow_b200.correlation import EEMFLOW_CDC_INDEX
    gen = torch.Generator().manual_seed(5)
    flow = torch.randn(1, 2, 6, 8, generator=gen, requires_grad=True)
    f1 = torch.randn(1, 16, 12, 16, generator=gen, requires_grad=True)
    f2 = torch.randn(1, 16, 12, 16, generator=gen, requires_grad=True)
    inter = (0.7 * torch.randn(1, 2, 12, 16, generator=gen)).requires_grad_(True)
    mask = torch.sigmoid(torch.randn(1, 1, 12, 16, generator=gen)).requires_grad_(True)

    def chain(mod, flow, f1, f2, inter, mask, on_gpu):
        if on_gpu:
            up = mod.upsample2d_flow_as(flow.clone(), f1, if_rate=True)
            w2 = mod.WarpingLayer_no_div()(f2, up)
            blended = mod.cdc_blend(up, inter, mask)
            f2w = mod.warp(f2, blended)
            cv = mod.correlation_select(f1, f2w, EEMFLOW_CDC_INDEX)
        else:
            up = ref_ops.upsample2d_flow_as(flow.clone(), f1, if_rate=True)
            w2 = ref_ops.warping_layer_no_div(f2, up)
            blended = ref_ops.cdc_blend(up, inter, mask)
            f2w = ref_ops.warp_exact(f2, blended)
            cv = ref_ops.correlation(f1, f2w, 4, index=EEMFLOW_CDC_INDEX)
        return (cv * cv).sum() + w2.sum()

    loss_ref = chain(None, flow, f1, f2, inter, mask, False)
    loss_ref.backward()
    leaves = [t.detach().cuda().requires_grad_(True) for t in (flow, f1, f2, inter, mask)]
    loss = chain(E, *leaves, True)
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    loss.backward()
    for name, a, b in zip(("flow", "f1", "f2", "inter", "mask"), leaves, (flow, f1, f2, inter, mask)):
        ok, e = close(a.grad, b.grad, 1e-3)
        assert ok, (name, e)


@pytest.mark.parametrize("B,D,H,W,levels,radius", [(2, 32, 8, 12, 3, 4), (1, 64, 9, 11, 3, 3), (2, 256, 16, 16, 4, 4), (1, 16, 5, 5, 2, 1)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_corrblock_backward(E, B, D, H, W, levels, radius, precision):
    """(Level sizes stay >= 2: a 1-wide level divides by W-1 = 0 in the reference's normalisation.)
    Gradient of two lookups (as ERAFT's iterations do) w.r.t. both feature maps vs torch autograd through
    the oracle's CorrBlock (matmul + avg_pool2d + grid_sample).  After an fp32 forward the backward GEMMs are exact
    fp32 (gate 2e-4); after a TF32 forward they run in TF32 on the tcgen05 kernel where the level's pitch allows
    (gate 3e-3 of the largest gradient, see test_batched_gemm_tf32)."""
    from eemflow_b200 import ops as O
    if precision == "tf32" and not O.tf32_supported(D, H, W):
        pytest.skip("shape not on the tcgen05 path")
    gen = torch.Generator().manual_seed(D + H)
    f1 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    f2 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    base = torch.stack(torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")[::-1], 0).float()[None]
    coords = [base + 2.0 * torch.randn(B, 2, H, W, generator=gen) for _ in range(2)]
    ref_pyr = ref_ops.corr_pyramid(f1, f2, levels)
    outs = [ref_ops.corr_lookup(ref_pyr, c, radius) for c in coords]
    gs = [torch.randn(o.shape, generator=gen) for o in outs]
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    a = f1.detach().cuda().requires_grad_(True)
    b = f2.detach().cuda().requires_grad_(True)
    blk = E.CorrBlock(a, b, num_levels=levels, radius=radius, precision=precision)
    assert all(lv.requires_grad for lv in blk.corr_pyramid if lv.numel())
    mine = [blk(c.cuda()) for c in coords]
    tol_fwd = 1e-5 if precision == "fp32" else 2e-3
    assert close(mine[0], outs[0], tol_fwd)[0]
    sum((o * g.cuda()).sum() for o, g in zip(mine, gs)).backward()
    tol_bwd = 2e-4 if precision == "fp32" else 3e-3
    ok1, e1 = close(a.grad, f1.grad, tol_bwd)
    ok2, e2 = close(b.grad, f2.grad, tol_bwd)
    assert ok1 and ok2, (e1, e2)


def test_corrblock_backward_partial_and_guards(E):
    gen = torch.Generator().manual_seed(3)
    f1 = torch.randn(1, 32, 16, 24, generator=gen)
    f2 = torch.randn(1, 32, 16, 24, generator=gen, requires_grad=True)
    coords = torch.stack(torch.meshgrid(torch.arange(16), torch.arange(24), indexing="ij")[::-1], 0).float()[None] + 0.3
    ref = ref_ops.corr_lookup(ref_ops.corr_pyramid(f1, f2, 4), coords, 4)
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    b = f2.detach().cuda().requires_grad_(True)
    blk = E.CorrBlock(f1.cuda(), b, precision="fp32")
    blk(coords.cuda()).backward(g.cuda())
    assert close(b.grad, f2.grad)[0]
    # CorrBlock.corr static method is differentiable too
    v = E.CorrBlock.corr(f1.cuda(), b, precision="fp32")
    assert v.requires_grad and v.shape == (1, 16, 24, 1, 16, 24)
    # coordinates that require grad take the differentiable composition (test_corrblock_coords_gradient)
    assert blk(coords.cuda().requires_grad_(True)).requires_grad
    # no gradient requested: plain inference path
    with torch.no_grad():
        assert not E.CorrBlock(f1.cuda(), b)(coords.cuda()).requires_grad


@pytest.mark.parametrize("batch,M,N,K", [(2, 64, 100, 37), (3, 256, 396, 1584), (1, 33, 65, 129)])
def test_batched_gemm_kernel_against_fp64(E, batch, M, N, K):
    """eem_batched_gemm_f32 (the pyramid backward's products, no cuBLAS): both operand orientations, alpha, accumulate;
    against an fp64 contraction, <= 2e-6 of the largest |result| (fp32 accumulation over K terms)."""
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(batch * 7 + K)
    A = torch.randn(batch, M, K, generator=gen).cuda()
    Bn = torch.randn(batch, K, N, generator=gen).cuda()
    Bt = Bn.transpose(1, 2).contiguous()
    ref = torch.bmm(A.double(), Bn.double())
    tol = 2e-6 * ref.abs().max().item() * max(1.0, (K / 256) ** 0.5)
    C = torch.empty(batch, M, N, device="cuda")
    ops.batched_gemm_(C, A, Bn, b_transposed=False, alpha=0.5)
    assert (C.double() - 0.5 * ref).abs().max().item() <= tol
    ops.batched_gemm_(C, A, Bt, b_transposed=True, alpha=0.25, accumulate=True)
    assert (C.double() - 0.75 * ref).abs().max().item() <= 2 * tol


def test_training_path_launches_no_library_gemm(E):
    """CorrBlock forward + backward under the profiler: every GEMM-class kernel is this library's (no cuBLAS / cutlass)."""
    from torch.profiler import ProfilerActivity, profile
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(1, 64, 16, 24, generator=gen).cuda().requires_grad_(True)
    b = torch.randn(1, 64, 16, 24, generator=gen).cuda().requires_grad_(True)
    coords = (ref_ops.coords_grid(1, 16, 24) + torch.randn(1, 2, 16, 24, generator=gen)).cuda()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = E.CorrBlock(a, b, num_levels=3, radius=4, precision="tf32")(coords)
        out.sum().backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert any("gemm_tf32_kernel" in n for n in names), names          # TF32 forward -> tcgen05 backward GEMMs
    assert not any(("cutlass" in n.lower()) or ("cublas" in n.lower()) or ("gemm" in n.lower() and "eem" not in n) for n in names), names


# ---- tcgen05 batched GEMM (csrc/gemm_tc.cu): the TF32 backward of the all-pairs pyramid ---------------------------
# Gate: <= 3e-3 of the largest |C| entry.  The tensor core truncates each operand to a 10-bit mantissa (up to 2^-10
# relative, always towards zero), so a product is low by up to 2e-3 and the bias does not average out in the largest
# entries: measured 0.7e-3 .. 1.0e-3 on B200 (profiles/r02/gemm_tc_diag.log).  A layout error gives O(1).
GEMM_TC_CASES = [
    (3, 256, 1584, 1584, True),    # d fmap1, level 0 of the MVSEC pyramid (K = 49.5 stages of 32)
    (2, 256, 1584, 396, True),
    (2, 256, 1584, 20, True),      # coarsest level: one ragged stage
    (3, 256, 1584, 1584, False),   # d fmap2, level 0 (dV read MN-major)
    (2, 256, 396, 1584, False),
    (2, 256, 20, 1584, False),     # a single ragged tile per sample
    (2, 64, 300, 200, True),
    (2, 128, 300, 200, False),
    (1, 32, 1000, 72, True),
    (5, 96, 132, 40, False),
]


@pytest.mark.parametrize("batch,M,N,K,bt", GEMM_TC_CASES)
def test_batched_gemm_tf32(E, batch, M, N, K, bt):
    from eemflow_b200 import ops as O
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(batch, M, K, generator=gen).cuda()
    B = torch.randn((batch, N, K) if bt else (batch, K, N), generator=gen).cuda()
    assert O.batched_gemm_tf32_supported(A, B, bt)
    launches = E._lib.lib().eem_launch_count()
    C = torch.full((batch, M, N), float("nan"), device="cuda")
    O.batched_gemm_(C, A, B, b_transposed=bt, alpha=0.5, precision="tf32")
    assert E._lib.lib().eem_launch_count() == launches + 1
    ref = 0.5 * torch.bmm(A.double(), B.double().transpose(1, 2) if bt else B.double())
    ok, e = close(C, ref, 3e-3)
    assert ok, e
    exact = torch.empty_like(C)
    O.batched_gemm_(exact, A, B, b_transposed=bt, alpha=0.5, precision="fp32")
    assert close(exact, ref, 1e-5)[0]
    O.batched_gemm_(C, A, B, b_transposed=bt, alpha=0.25, accumulate=True, precision="tf32")
    ok, e = close(C, 1.5 * ref, 3e-3)
    assert ok, e


def test_batched_gemm_tf32_shape_rules(E):
    """Shapes TMA cannot address are answered by `supported` (and take the FFMA kernel inside batched_gemm_); the
    C entry point itself refuses them loudly."""
    from eemflow_b200 import ops as O
    lib = E._lib.lib()
    assert lib.eem_batched_gemm_tf32_supported(2, 256, 1584, 1584, 1584, 1584, 256 * 1584, 1584 * 1584, 1) == 1
    assert lib.eem_batched_gemm_tf32_supported(2, 256, 1584, 99, 99, 99, 256 * 99, 1584 * 99, 1) == 0      # 9x11 level
    assert lib.eem_batched_gemm_tf32_supported(2, 48, 1584, 396, 396, 396, 48 * 396, 1584 * 396, 1) == 0   # M % 32
    assert lib.eem_batched_gemm_tf32_supported(2, 288, 1584, 396, 396, 396, 288 * 396, 1584 * 396, 1) == 0 # M > 256
    A = torch.randn(2, 256, 99).cuda()
    B = torch.randn(2, 64, 99).cuda()
    C = torch.empty(2, 256, 64).cuda()
    with pytest.raises(NotImplementedError):
        E._lib.check(lib.eem_batched_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), 2, 256, 64, 99, 99, 99, 64, 256 * 99,
                                               64 * 99, 256 * 64, 1, 1.0, 0, None))
    O.batched_gemm_(C, A, B, b_transposed=True, precision="tf32")          # falls back to the exact kernel
    assert close(C, torch.bmm(A.double(), B.double().transpose(1, 2)), 1e-5)[0]


@pytest.mark.parametrize("B,D,H,W,levels", [(2, 256, 16, 16, 4), (2, 64, 36, 44, 4), (1, 128, 12, 20, 3)])
def test_corrblock_backward_tf32_gemms(E, monkeypatch, B, D, H, W, levels):
    """CorrBlock backward with the pyramid GEMMs on the tcgen05 kernel (levels whose pitch TMA cannot address -- 9x11
    of 36x44 -- take the FFMA kernel) against torch autograd through the oracle.  Gate 3e-3 of the largest gradient."""
    monkeypatch.setenv("EEMFLOW_B200_CORR_BACKWARD", "tf32")
    gen = torch.Generator().manual_seed(D + H)
    f1 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    f2 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    base = torch.stack(torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")[::-1], 0).float()[None]
    coords = [base + 2.0 * torch.randn(B, 2, H, W, generator=gen) for _ in range(2)]
    ref_pyr = ref_ops.corr_pyramid(f1, f2, levels)
    outs = [ref_ops.corr_lookup(ref_pyr, c, 4) for c in coords]
    gs = [torch.randn(o.shape, generator=gen) for o in outs]
    sum((o * g).sum() for o, g in zip(outs, gs)).backward()
    a = f1.detach().cuda().requires_grad_(True)
    b = f2.detach().cuda().requires_grad_(True)
    blk = E.CorrBlock(a, b, num_levels=levels, radius=4, precision="tf32")
    mine = [blk(c.cuda()) for c in coords]
    sum((o * g.cuda()).sum() for o, g in zip(mine, gs)).backward()
    ok1, e1 = close(a.grad, f1.grad, 3e-3)
    ok2, e2 = close(b.grad, f2.grad, 3e-3)
    assert ok1 and ok2, (e1, e2)


@pytest.mark.parametrize("N,C,H,W,Ho,Wo", [(2, 3, 9, 13, 7, 5), (6, 1, 12, 16, 9, 9), (1, 8, 5, 4, 33, 3)])
def test_bilinear_sampler_backward(E, N, C, H, W, Ho, Wo):
    """bilinear_sampler (model/model_utils.py:7-21) under autograd: gradients to the image and to the coordinates vs
    torch autograd through the oracle's F.grid_sample composition; samples outside the map included."""
    gen = torch.Generator().manual_seed(N * H + W)
    img = torch.randn(N, C, H, W, generator=gen, requires_grad=True)
    coords = torch.stack([torch.rand(N, Ho, Wo, generator=gen) * (W + 3) - 2, torch.rand(N, Ho, Wo, generator=gen) * (H + 3) - 2], -1)
    coords.requires_grad_(True)
    ref, ref_mask = ref_ops.bilinear_sampler(img, coords, mask=True)
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    a = img.detach().cuda().requires_grad_(True)
    c = coords.detach().cuda().requires_grad_(True)
    out, mask = E.bilinear_sampler(a, c, mask=True)
    assert out.requires_grad and not mask.requires_grad
    assert close(out, ref, 1e-5)[0] and torch.equal(mask.cpu(), ref_mask.float())
    out.backward(g.cuda())
    ok1, e1 = close(a.grad, img.grad)
    ok2, e2 = close(c.grad, coords.grad)
    assert ok1 and ok2, (e1, e2)
    # one-sided: only the coordinates need a gradient
    c2 = coords.detach().cuda().requires_grad_(True)
    E.bilinear_sampler(img.detach().cuda(), c2).backward(g.cuda())
    assert close(c2.grad, coords.grad)[0]


def test_corrblock_coords_gradient(E):
    """CorrBlock.__call__ with coordinates that require grad (the reference differentiates them through grid_sample;
    its shipped callers detach them, model/eraft.py:141): value and all three gradients vs the oracle."""
    gen = torch.Generator().manual_seed(11)
    B, D, H, W, levels, radius = 2, 32, 8, 12, 3, 3
    f1 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    f2 = torch.randn(B, D, H, W, generator=gen, requires_grad=True)
    base = torch.stack(torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")[::-1], 0).float()[None]
    coords = (base + 1.5 * torch.randn(B, 2, H, W, generator=gen) + 0.25).requires_grad_(True)
    ref = ref_ops.corr_lookup(ref_ops.corr_pyramid(f1, f2, levels), coords, radius)
    g = torch.randn(ref.shape, generator=gen)
    ref.backward(g)
    a = f1.detach().cuda().requires_grad_(True)
    b = f2.detach().cuda().requires_grad_(True)
    c = coords.detach().cuda().requires_grad_(True)
    blk = E.CorrBlock(a, b, num_levels=levels, radius=radius, precision="fp32")
    out = blk(c)
    assert close(out, ref, 1e-5)[0]
    with torch.no_grad():                                   # the same values as the fused lookup kernel
        assert close(out, blk(c.detach()), 1e-5)[0]
    out.backward(g.cuda())
    for name, mine, theirs in (("coords", c.grad, coords.grad), ("fmap1", a.grad, f1.grad), ("fmap2", b.grad, f2.grad)):
        ok, e = close(mine, theirs)
        assert ok, (name, e)
    # pyramid without grad, coordinates with: the gradient still arrives (round 1 returned a detached result)
    c2 = coords.detach().cuda().requires_grad_(True)
    E.CorrBlock(f1.detach().cuda(), f2.detach().cuda(), num_levels=levels, radius=radius, precision="fp32")(c2).backward(g.cuda())
    assert close(c2.grad, coords.grad)[0]


def test_batched_gemm_tf32_multi_segments(E):
    """One launch, K loop over several segments (the pyramid's levels): C = alpha * sum_s A_s . B_s^T (+ C)."""
    from eemflow_b200 import ops as O
    gen = torch.Generator().manual_seed(9)
    batch, M, N = 3, 256, 1000
    Ks = (1584, 396, 100, 20)
    As = [torch.randn(batch, M, K, generator=gen).cuda() for K in Ks]
    Bs = [torch.randn(batch, N, K, generator=gen).cuda() for K in Ks]
    ref = sum(torch.bmm(a.double(), b.double().transpose(1, 2)) for a, b in zip(As, Bs))
    launches = E._lib.lib().eem_launch_count()
    C = torch.full((batch, M, N), float("nan"), device="cuda")
    O.batched_gemm_tf32_multi_(C, As, Bs, b_transposed=True, alpha=0.5)
    assert E._lib.lib().eem_launch_count() == launches + 1
    ok, e = close(C, 0.5 * ref, 3e-3)
    assert ok, e
    O.batched_gemm_tf32_multi_(C, As[1:3], Bs[1:3], b_transposed=True, alpha=1.0, accumulate=True)
    ref2 = 0.5 * ref + sum(torch.bmm(a.double(), b.double().transpose(1, 2)) for a, b in zip(As[1:3], Bs[1:3]))
    ok, e = close(C, ref2, 3e-3)
    assert ok, e
    Bn = [b.transpose(1, 2).contiguous() for b in Bs]                    # the MN-major operand form
    O.batched_gemm_tf32_multi_(C, As, Bn, b_transposed=False, alpha=0.5)
    ok, e = close(C, 0.5 * ref, 3e-3)
    assert ok, e
    with pytest.raises(NotImplementedError):
        O.batched_gemm_tf32_multi_(C, [torch.randn(batch, M, 99).cuda()], [torch.randn(batch, N, 99).cuda()], b_transposed=True)
