"""GPU parity of the loader / evaluation helpers (csrc/eval_ops.cu) against the goldens produced by executing the
reference's own source (tests/golden/eval.npz, oracle/gen_golden.py::gen_eval) and against the CPU oracle.
Gates: masks, counts, mesh flows and the bin sum are bit-exact; the EPE sums are fp32 numbers accumulated in a
different order (double on the device), gate 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import ref_ops
from tests.test_oracle_golden import _hashed_flow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import eemflow_b200
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    return eemflow_b200


@pytest.fixture(scope="module")
def g(golden):
    return golden("eval")


@pytest.mark.parametrize("name,kind,is_car", [("sparse", "sparse", False), ("dense", "dense", False), ("car", "sparse", True)])
def test_flow_error_matches_reference(E, g, name, kind, is_car):
    gt, pred, ev = (torch.from_numpy(g[f"fe_{name}_{k}"]) for k in ("gt", "pred", "ev"))
    res = E.flow_error(gt.cuda(), pred.cuda(), ev.cuda(), is_car, kind)
    ref = g[f"fe_{name}_res"]
    mine = np.array([float(v) for v in res])
    assert mine[3] == ref[3] and mine[3] > 100                         # n_points
    assert mine[1] == ref[1] and mine[2] == ref[2]                     # the two percentages (counts are exact)
    for k in (0, 4, 5, 6):
        assert abs(mine[k] - ref[k]) <= 1e-5 * abs(ref[k]), (k, mine[k], ref[k])
    # CPU tensors are accepted as well (uploaded), like the reference's call site passes them
    res2 = E.flow_error(gt, pred, ev, is_car, kind)
    assert [float(v) for v in res2] == [float(v) for v in res]


def test_flow_error_empty_and_batch(E):
    gt = torch.zeros(2, 2, 8, 8)
    pred = torch.ones(2, 2, 8, 8)
    assert E.flow_error(gt, pred, torch.ones(1, 1, 8, 8)) == (0, 0.0, 0.0, 0, 0, 0, 0)
    from eemflow_b200.eval_utils import flow_error_stats
    gen = torch.Generator().manual_seed(0)
    gt = torch.randn(3, 2, 20, 30, generator=gen)
    pred = gt + 0.5 * torch.randn(3, 2, 20, 30, generator=gen)
    ev = (torch.rand(3, 1, 20, 30, generator=gen) < 0.5).float()
    stats = flow_error_stats(gt.cuda(), pred.cuda(), ev.cuda()).cpu()
    for b in range(3):
        ref = ref_ops.flow_error(gt[b:b + 1], pred[b:b + 1], ev[b:b + 1], False, "sparse")
        assert int(stats[b, 0]) == ref[3]
        assert abs(stats[b, 3].item() - float(ref[4])) <= 1e-5 * float(ref[4])


@pytest.mark.parametrize("name", ["hrem", "small"])
def test_motion_propagate_matches_reference(E, g, name):
    h, w = (int(v) for v in g[f"mp_{name}_hw"])
    ff = _hashed_flow(h, w)
    xm, ym = E.motion_propagate(ff, h, w)
    assert xm.dtype == np.float64 and xm.shape == (16, 16)
    assert np.array_equal(xm, g[f"mp_{name}_x"]) and np.array_equal(ym, g[f"mp_{name}_y"])


def test_motion_propagate_batch_other_meshes(E):
    from eemflow_b200.eval_utils import motion_propagate_batch
    rng = np.random.default_rng(4)
    ff = rng.normal(0, 3, size=(3, 90, 120, 2)).astype(np.float32)
    for mesh, radius in ((16, 3), (8, 2), (5, 1), (16, 0)):
        out = motion_propagate_batch(torch.from_numpy(ff).cuda(), mesh, radius).cpu().numpy()
        for b in range(3):
            xm, ym = ref_ops.motion_propagate(ff[b], 90, 120, mesh, radius)
            assert np.array_equal(out[b, 0].astype(np.float64), xm) and np.array_equal(out[b, 1].astype(np.float64), ym), (mesh, radius, b)


def test_event_mask_and_event_valid(E, g):
    h, w = (int(v) for v in g["mask_hw"])
    seq = E.EventSequence(None, {"height": h, "width": w}, features=g["mask_events"].copy())
    m = E.event_mask(seq)
    assert m.dtype == torch.bool and tuple(m.shape) == (1, h, w)
    assert np.array_equal(m[0].cpu().numpy(), g["mask_out"])
    s = E.event_valid_from_volume(torch.from_numpy(g["binsum_in"]).cuda())
    assert tuple(s.shape) == (1, h, w)
    assert np.array_equal(s[0].cpu().numpy(), g["binsum_out"])
    # batched windows
    from eemflow_b200.eval_utils import event_mask_batch
    rng = np.random.default_rng(9)
    evs = [np.stack([np.sort(rng.uniform(0, 1, n)), rng.uniform(-2, w + 2, n), rng.uniform(-2, h + 2, n), np.ones(n)], 1) for n in (500, 1, 3000)]
    off = torch.tensor([0, 500, 501, 3501], dtype=torch.int64).cuda()
    mb = event_mask_batch(torch.from_numpy(np.concatenate(evs)).cuda(), off, 3000, h, w).cpu().numpy()
    for k, ev in enumerate(evs):
        assert np.array_equal(mb[k, 0], ref_ops.event_mask(ev, h, w)), k
