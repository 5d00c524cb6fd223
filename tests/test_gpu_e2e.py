"""GPU end-to-end parity (BASELINE config 0): synthetic events -> 5-bin voxel grids -> EEMFlow_cdc forward,
against the REAL reference model run on the CPU (tests/golden/e2e_eemflow_cdc.npz, oracle/gen_golden.py).
The drop-in model shares parameter names/shapes/order with the reference; both sides use the RNG-free
parameter values of oracle/det_weights.py.  Gate: EPE-relative (mean |dflow| / mean |flow|) <= 1e-3 per
prediction with cuDNN TF32 disabled, as the north star states for correlation and flow outputs."""
import numpy as np
import pytest
import torch

from oracle.det_weights import set_deterministic_weights

pytestmark = pytest.mark.gpu


class Seq:
    def __init__(self, features, h, w):
        self.features, self.image_height, self.image_width = features, h, w


@pytest.mark.parametrize("local_corr", ["fp32", "tf32"])
def test_eemflow_cdc_forward_from_events(golden, local_corr):
    """local_corr="tf32": the five local 9x9 correlations on the tcgen05 banded-GEMM kernel (TF32 products, fp32
    accumulation) -- same end-to-end gate as the north star states for the TF32 correlation volume."""
    import eemflow_b200 as E
    from eemflow_b200.models import EEMFlow_cdc
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    g = golden("e2e_eemflow_cdc")
    nb, h, w = (int(v) for v in g["shape"])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        E.set_local_corr_precision(local_corr)
        enc = E.EventSequenceToVoxelGrid_Pytorch(nb, gpu=True, normalize=True, forkserver=False)
        v1 = enc(Seq(g["events1"].copy(), h, w))[None]
        v2 = enc(Seq(g["events2"].copy(), h, w))[None]
        assert np.abs(v1.cpu().numpy() - g["voxel1"]).max() <= 1e-5 * max(1.0, np.abs(g["voxel1"]).max())
        net = EEMFlow_cdc(None, groups=3, n_first_channels=nb)
        set_deterministic_weights(net)
        net = net.cuda().eval()
        net.change_imagesize((h, w))
        with torch.no_grad():
            (e1, e2), flows = net(events1=v1, events2=v2)
        assert e1 is v1 and e2 is v2 and len(flows) == 5
        for k, f in enumerate(flows):
            ref = g[f"flow{k}"]
            assert tuple(f.shape) == ref.shape == (1, 2, h, w)
            d = np.abs(f.cpu().numpy() - ref)
            rel = d.mean() / np.abs(ref).mean()
            assert rel <= 1e-3, (local_corr, k, rel, d.max())
    finally:
        E.set_local_corr_precision(None)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_eemflow_cdc_trains_through_the_kernels():
    """One optimisation step through the drop-in model: gradients flow through the CUDA backward kernels."""
    from eemflow_b200.models import EEMFlow_cdc
    torch.manual_seed(0)
    net = EEMFlow_cdc(None, groups=3, n_first_channels=5).cuda().train()
    net.change_imagesize((64, 96))
    v1 = torch.randn(2, 5, 64, 96, device="cuda")
    v2 = torch.randn(2, 5, 64, 96, device="cuda")
    target = torch.zeros(2, 2, 64, 96, device="cuda")
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    losses = []
    for _ in range(3):
        _, flows = net(events1=v1, events2=v2)
        loss = sum((f - target).abs().mean() for f in flows)   # L1 sequence loss, train_mvsec.py:201-227
        opt.zero_grad()
        loss.backward()
        grads = [p.grad for p in net.parameters() if p.grad is not None]
        assert len(grads) > 100 and all(torch.isfinite(g_).all() for g_ in grads)
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all()


@pytest.mark.parametrize("precision,gate", [("fp32", 2e-4), ("tf32", 1e-3), ("tf32_f16", 1e-3)])
def test_eraft_forward_from_events(golden, precision, gate):
    """ERAFT, the caller of CorrBlock: events -> voxel grids -> encoders -> all-pairs pyramid -> 12 x (lookup + GRU
    update) -> convex upsampling, against the REAL reference ERAFT on the CPU (tests/golden/e2e_eraft.npz).
    Any error of the correlation volume is fed back through 12 recurrent updates, which is why the north star
    states the TF32 tolerance end to end: EPE-relative <= 1e-3.  (The golden's weights make the flow depend
    on the lookups: zeroing them changes it by 90 %, see oracle/gen_golden.py::gen_e2e_eraft.)"""
    import eemflow_b200 as E
    from eemflow_b200.models import ERAFT
    from oracle.det_weights import set_hashed_weights
    g = golden("e2e_eraft")
    nb, h, w = (int(v) for v in g["shape"])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        enc = E.EventSequenceToVoxelGrid_Pytorch(nb, gpu=True, normalize=True, forkserver=False)
        v1 = enc(Seq(g["events1"].copy(), h, w))[None]
        v2 = enc(Seq(g["events2"].copy(), h, w))[None]
        net = ERAFT(None, n_first_channels=nb, corr_precision=precision)
        set_hashed_weights(net, weight_gain=0.8)
        net = net.cuda().eval()
        net.change_imagesize((h, w))
        with torch.no_grad():
            _, flows = net(events1=v1, events2=v2, iters=12)
        assert len(flows) == 12
        for k in (0, 5, 11):
            ref = g[f"flow{k}"]
            f = flows[k].cpu().numpy()
            assert f.shape == ref.shape == (1, 2, h, w)
            rel = np.abs(f - ref).mean() / np.abs(ref).mean()
            assert rel <= gate, (precision, k, rel)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_eraft_trains_through_the_kernels():
    """Gradients reach the feature encoder through the CorrBlock backward (lookup gradient + pyramid GEMMs)."""
    from eemflow_b200.models import ERAFT
    torch.manual_seed(0)
    net = ERAFT(None, n_first_channels=5).cuda().train()
    net.change_imagesize((128, 192))      # 1/8 maps of 16x24: every pyramid level stays >= 2 wide (a 1-wide level
    v1 = torch.randn(2, 5, 128, 192, device="cuda")   # is NaN in the reference too: division by W-1 = 0)
    v2 = torch.randn(2, 5, 128, 192, device="cuda")
    _, flows = net(events1=v1, events2=v2, iters=3)
    loss = sum(f.abs().mean() for f in flows)
    loss.backward()
    for p in (net.fnet.conv1.weight, net.fnet.conv2.weight, net.update_block.encoder.convc1.weight, net.cnet.conv1.weight):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0


@pytest.mark.parametrize("which", ["eemflow_cdc", "eraft"])
def test_graphed_inference_matches_eager(which):
    """One-CUDA-graph replay of a model forward gives the eager result bit for bit (same kernels, same order)."""
    from eemflow_b200.models import EEMFlow_cdc, ERAFT, GraphedInference
    torch.manual_seed(1)
    net = (EEMFlow_cdc(None, groups=3, n_first_channels=5) if which == "eemflow_cdc" else ERAFT(None, n_first_channels=5)).cuda().eval()
    kw = {} if which == "eemflow_cdc" else {"iters": 4}
    h, w = 128, 192
    net.change_imagesize((h, w))
    fast = GraphedInference(net)
    for trial in range(3):                      # first call captures, later calls replay with new inputs
        v1 = torch.randn(2, 5, h, w, device="cuda")
        v2 = torch.randn(2, 5, h, w, device="cuda")
        with torch.no_grad():
            _, ref = net(events1=v1, events2=v2, **kw)
        _, got = fast(v1, v2, **kw)
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert torch.equal(a, b), (which, trial, (a - b).abs().max().item())
    assert len(fast._captured) == 1
