"""GPU parity: voxelization kernels (csrc/voxelize.cu) against the golden vectors of the real
reference and against the CPU oracle on seeded inputs.  Run with `-m gpu` on a B200.

Gates: vote addresses bit-exact; raw grid bit-exact in deterministic mode and
<= 1e-5 relative (|a-b| <= 1e-5*max(|b|,1)) in atomic mode; normalised grid <= 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_ops
from tests.conftest import cases_of

pytestmark = pytest.mark.gpu


def rel_close(a, b, tol=1e-5):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0)


class Seq:
    def __init__(self, features, h, w):
        self.features, self.image_height, self.image_width = features, h, w


def make_events(rng, n, h, w, clustered=False):
    t = np.sort(rng.uniform(0.0, 0.05, size=n)) * 1e6
    t -= t[0]
    if clustered:  # 80 % of the events in a few Gaussian blobs (~5 % of the pixels)
        k = int(0.8 * n)
        centres = rng.uniform([0, 0], [w, h], size=(8, 2))
        which = rng.integers(0, 8, size=k)
        sig = 0.035 * min(h, w)
        xy = centres[which] + rng.normal(0, sig, size=(k, 2))
        x = np.concatenate([np.clip(np.round(xy[:, 0]), 0, w - 1), rng.integers(0, w, size=n - k)])
        y = np.concatenate([np.clip(np.round(xy[:, 1]), 0, h - 1), rng.integers(0, h, size=n - k)])
        perm = rng.permutation(n)
        x, y = x[perm], y[perm]
    else:
        x = rng.integers(0, w, size=n).astype(np.float64)
        y = rng.integers(0, h, size=n).astype(np.float64)
    p = 2.0 * rng.integers(0, 2, size=n) - 1.0
    return np.stack([t, x.astype(np.float64), y.astype(np.float64), p], axis=1)


@pytest.fixture(scope="module")
def V():
    import eemflow_b200
    assert torch.cuda.is_available(), "GPU tests selected but no CUDA device is visible"
    return eemflow_b200.EventSequenceToVoxelGrid_Pytorch


def test_golden_raw_atomic_and_deterministic(golden, V):
    g = golden("voxel")
    for name in cases_of(g):
        nb, h, w = (int(v) for v in g[f"{name}__shape"])
        ev = g[f"{name}__events"]
        ref = g[f"{name}__raw"]
        out_a = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev.copy(), h, w)).cpu().numpy()
        assert out_a.shape == ref.shape and out_a.dtype == np.float32
        assert rel_close(out_a, ref).all(), (name, np.abs(out_a - ref).max())
        out_d = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev.copy(), h, w)).cpu().numpy()
        assert np.array_equal(out_d, ref), (name, np.abs(out_d - ref).max())     # bit-exact


def test_golden_normalized(golden, V):
    g = golden("voxel")
    for name in cases_of(g):
        nb, h, w = (int(v) for v in g[f"{name}__shape"])
        ev = g[f"{name}__events"]
        ref = g[f"{name}__norm"]
        for det in (False, True):
            out = V(nb, gpu=True, normalize=True, forkserver=False, deterministic=det)(Seq(ev.copy(), h, w)).cpu().numpy()
            assert rel_close(out, ref).all(), (name, det, np.abs(out - ref).max())


@pytest.mark.parametrize("n,nb,h,w,clustered", [
    (30_000, 5, 260, 346, False),      # MVSEC dt1 window
    (120_000, 5, 260, 346, True),      # MVSEC dt4, clustered
    (1_000_000, 15, 720, 1280, False),  # HREM-shaped, 1M events
    (600_000, 15, 720, 1280, True),
])
def test_oracle_parity(V, n, nb, h, w, clustered):
    rng = np.random.default_rng(n + nb)
    ev = make_events(rng, n, h, w, clustered)
    ref_raw, dropped, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)
    assert dropped == 0
    out_d = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev, h, w)).cpu().numpy()
    assert np.array_equal(out_d, ref_raw)                                       # bit-exact
    out_a = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert rel_close(out_a, ref_raw).all(), np.abs(out_a - ref_raw).max()
    ref_norm = ref_ops.voxelize(ev, nb, h, w, normalize=True).numpy()
    out_n = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert rel_close(out_n, ref_norm).all(), np.abs(out_n - ref_norm).max()


@pytest.mark.parametrize("n_windows,shape", [(40, (5, 64, 96)), (3, (5, 260, 346)), (2, (15, 60, 62)), (1, (1, 1, 4))])
def test_normalize_cluster_path_matches_split_path(n_windows, shape):
    """K2 as ONE launch per call -- a cluster of 8 CTAs per window, statistics exchanged over distributed shared memory:
    the streaming form (default: two passes over the slice, every window in flight) and the shared-memory-resident form
    (EEM_VOXEL_NORM=cluster) -- against the two-kernel path (EEM_VOXEL_NORM=split) and the oracle's formula: non-zero count exact, mean / std and the
    normalised voxels <= 1e-6 relative (the two paths add the same fp64 terms in a different order).  40 windows: more
    than the resident clusters, so clusters loop; all-zero, one-non-zero (std = NaN -> v - mean) and dense windows."""
    import os
    from eemflow_b200 import ops
    gen = torch.Generator().manual_seed(n_windows)
    grid = torch.randn((n_windows,) + shape, generator=gen)
    grid[torch.rand(grid.shape, generator=gen) < 0.7] = 0.0
    if n_windows >= 3:
        grid[1].zero_()                                   # no events at all
        grid[2].zero_()
        grid[2].view(-1)[3] = 2.5                         # a single non-zero voxel
    outs, stats = {}, {}
    for mode in ("stream", "cluster", "split"):
        os.environ["EEM_VOXEL_NORM"] = mode
        try:
            g = grid.cuda()
            st = torch.zeros(n_windows, 3, dtype=torch.float64, device="cuda")
            ops.voxel_normalize_(g, st)
            outs[mode], stats[mode] = g.cpu(), st.cpu()
        finally:
            os.environ.pop("EEM_VOXEL_NORM", None)
    assert torch.equal(stats["cluster"][:, 0], stats["split"][:, 0]) and torch.equal(stats["stream"][:, 0], stats["split"][:, 0])
    assert torch.equal(torch.nan_to_num(outs["stream"], nan=-7.0), torch.nan_to_num(outs["cluster"], nan=-7.0))   # same arithmetic, same order
    assert torch.equal(stats["cluster"][:, 0], (grid.flatten(1) != 0).sum(1).double())
    for k in range(n_windows):
        nz = grid[k][grid[k] != 0].double()
        if nz.numel() > 1:
            mean, std = nz.mean().item(), nz.std().item()
            assert abs(stats["cluster"][k, 1].item() - mean) <= 1e-6 * max(1.0, abs(mean))
            assert abs(stats["cluster"][k, 2].item() - std) <= 1e-6 * std
    a, b = outs["cluster"], outs["split"]
    assert torch.equal(a == 0, b == 0) and torch.equal(torch.isnan(a), torch.isnan(b))
    assert ((a - b).abs() <= 1e-6 * b.abs().clamp_min(1.0)).all()
    if n_windows >= 3:
        assert a[1].abs().max().item() == 0.0 and a[2].view(-1)[3].item() == 0.0     # 2.5 - mean(2.5)


def test_cluster_path_fixed_point_and_overflow_redo(V):
    """The opt-in cluster-resident voxelizer (EEM_VOXEL_PATH=cluster; measured slower than the L2-atomic kernels, kept as a
    tested experiment): votes are integer adds (2^-22 units) into the distributed shared memory of 8 CTAs.  (1) ordinary
    windows: <= 1e-5 relative against the C oracle and against the default L2-atomic kernels; (2) a hot pixel collecting
    ~1500 same-sign votes per bin exceeds half the int32 range: the cluster must notice and redo the window with float
    adds (gate scaled by the vote mass like every order-free sum); (3) a polarity of 100 is outside the fixed-point
    vote range and takes the same redo."""
    import os
    rng = np.random.default_rng(77)
    nb, h, w = 5, 260, 346
    enc = V(nb, gpu=True, normalize=False, forkserver=False)
    ev = make_events(rng, 30_000, h, w)
    ref, _, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)
    out_l2 = enc(Seq(ev, h, w)).cpu().numpy()
    os.environ["EEM_VOXEL_PATH"] = "cluster"
    try:
        _cluster_path_checks(V, enc, ev, ref, out_l2, rng, nb, h, w)
    finally:
        os.environ.pop("EEM_VOXEL_PATH", None)


def _cluster_path_checks(V, enc, ev, ref, out_l2, rng, nb, h, w):
    out = enc(Seq(ev, h, w)).cpu().numpy()
    assert rel_close(out, ref).all(), np.abs(out - ref).max()
    assert rel_close(out, out_l2).all()
    # (2) hot pixel: 6000 events, all polarity +1, at (x, y) = (17, 33), inside a window of uniform events
    hot = make_events(rng, 26_000, h, w)
    hot[::4, 1], hot[::4, 2], hot[::4, 3] = 17.0, 33.0, 1.0
    ref, _, _ = c_oracle.voxelize(hot, nb, h, w, normalize=False)
    assert np.abs(ref).max() > 600.0                       # far beyond the +-256 the fixed-point cells may reach
    mass = _vote_mass(hot, nb, h, w)
    batch = enc.voxelize_batch([Seq(ev, h, w), Seq(hot, h, w), Seq(ev, h, w)]).cpu().numpy()
    assert order_free_close(batch[1], ref, mass).all(), np.abs(batch[1] - ref).max()
    assert rel_close(batch[0], out).all()
    assert rel_close(batch[2], batch[0]).all()
    # normalised: same window through the fused statistics
    ref_n = ref_ops.voxelize(hot, nb, h, w, normalize=True).numpy()
    out_n = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(hot, h, w)).cpu().numpy()
    assert np.abs(out_n - ref_n).max() <= 1e-4 * np.abs(ref_n).max()
    # (3) a large polarity value
    big = ev.copy()
    big[5, 3] = 100.0
    ref, _, _ = c_oracle.voxelize(big, nb, h, w, normalize=False)
    out = enc(Seq(big, h, w)).cpu().numpy()
    assert rel_close(out, ref).all(), np.abs(out - ref).max()


def test_batched_windows_ragged(V):
    rng = np.random.default_rng(5)
    h, w, nb = 64, 96, 5
    seqs = [Seq(make_events(rng, n, h, w), h, w) for n in (1, 2, 777, 5000, 31, 12345)]
    enc = V(nb, gpu=True, normalize=True, forkserver=False)
    batch = enc.voxelize_batch(seqs).cpu().numpy()
    assert batch.shape == (len(seqs), nb, h, w)
    for k, s in enumerate(seqs):
        ref = ref_ops.voxelize(s.features, nb, h, w, normalize=True).numpy()
        assert rel_close(batch[k], ref).all(), k
    det = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_batch(seqs).cpu().numpy()
    for k, s in enumerate(seqs):
        ref, _, _ = c_oracle.voxelize(s.features, nb, h, w, normalize=False)
        assert np.array_equal(det[k], ref), k


def _full_size_case(n, seed, clustered=False):
    rng = np.random.default_rng(seed)
    nb, h, w = 15, 720, 1280
    ev = make_events(rng, n, h, w, clustered)
    ref_raw, dropped, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)      # per-voxel CPU oracle (plain C)
    assert dropped == 0
    return ev, ref_raw, nb, h, w


def _vote_mass(ev, nb, h, w):
    """sum_i |vote_i| per voxel (the oracle run on the same events with every polarity set to +1)."""
    pos = ev.copy()
    pos[:, 3] = 1.0
    return c_oracle.voxelize(pos, nb, h, w, normalize=False)[0]


def order_free_close(a, b, mass, tol=1e-5):
    """Gate of the ORDER-FREE (atomic) mode: |a - b| <= 1e-5 * max(|b|, 1, sum_i |vote_i|).  On voxels that collect
    hundreds of +-1 votes (clustered streams) the reference's own fp32 sum depends on the order of its adds by
    ~sqrt(n) * 2^-24 * sum|vote|, so an order-free sum can only be held to the summation forward-error scale; for
    voxels with a few votes this is the plain 1e-5 * max(|b|, 1) gate.  The deterministic mode is held to bit-exactness."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= tol * np.maximum(np.maximum(np.abs(b), 1.0), mass)


@pytest.mark.parametrize("clustered", [False, True])
def test_full_size_hrem_dt1_per_voxel(V, clustered):
    """BASELINE configs[2] size (10 M events, 15x720x1280), EVERY voxel against the C oracle: deterministic mode
    bit-exact, atomic mode <= 1e-5 relative, normalised <= 1e-5 relative; rows and packed columns."""
    ev, ref_raw, nb, h, w = _full_size_case(10_000_000, 9 + int(clustered), clustered)
    mass = _vote_mass(ev, nb, h, w)
    out_d = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev, h, w)).cpu().numpy()
    assert np.array_equal(out_d, ref_raw)                                                         # bit-exact
    del out_d
    out_a = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert order_free_close(out_a, ref_raw, mass).all(), np.abs(out_a - ref_raw).max()
    if not clustered:
        assert rel_close(out_a, ref_raw).all(), np.abs(out_a - ref_raw).max()
    del out_a
    cols = [{"t": np.ascontiguousarray(ev[:, 0]), "x": ev[:, 1].astype(np.int16), "y": ev[:, 2].astype(np.int16),
             "p": ev[:, 3].astype(np.int8)}]
    out_c = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_columns(cols, h, w)[0].cpu().numpy()
    assert np.array_equal(out_c, ref_raw)                                                         # packed columns, bit-exact
    out_ca = V(nb, gpu=True, normalize=False, forkserver=False).voxelize_columns(cols, h, w)[0].cpu().numpy()
    assert order_free_close(out_ca, ref_raw, mass).all(), np.abs(out_ca - ref_raw).max()
    del out_c, out_ca
    ref_norm, _, stats = c_oracle.voxelize(ev, nb, h, w, normalize=True)
    out_n = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert order_free_close(out_n, ref_norm, mass / max(float(stats[2]), 1e-30)).all(), np.abs(out_n - ref_norm).max()


def test_full_size_hrem_dt4_per_voxel(V):
    """BASELINE configs[3] size (40 M events per window: the auto-selected path for event-dominated windows),
    every voxel against the C oracle: atomic <= 1e-5 relative, deterministic bit-exact, normalised <= 1e-5."""
    ev, ref_raw, nb, h, w = _full_size_case(40_000_000, 19)
    out_a = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert rel_close(out_a, ref_raw).all(), np.abs(out_a - ref_raw).max()      # uniform stream: ~3 votes per voxel
    del out_a
    out_d = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev, h, w)).cpu().numpy()
    assert np.array_equal(out_d, ref_raw)
    del out_d
    ref_norm, _, _ = c_oracle.voxelize(ev, nb, h, w, normalize=True)
    out_n = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
    assert rel_close(out_n, ref_norm).all(), np.abs(out_n - ref_norm).max()


def test_device_contract_and_errors(V):
    rng = np.random.default_rng(3)
    ev = make_events(rng, 1000, 32, 48)
    out_cpu = V(5, gpu=False, normalize=True, forkserver=False)(Seq(ev, 32, 48))
    assert out_cpu.device.type == "cpu" and out_cpu.dtype == torch.float32 and tuple(out_cpu.shape) == (5, 32, 48)
    out_gpu = V(5, gpu=True, normalize=True, forkserver=False)(Seq(ev, 32, 48))
    assert out_gpu.is_cuda
    assert torch.equal(out_gpu.cpu() != 0, out_cpu != 0)
    ev_before = ev.copy()
    V(5, gpu=True, forkserver=False)(Seq(ev, 32, 48))
    assert np.array_equal(ev, ev_before)                          # features are not modified
    with pytest.raises(AssertionError):
        V(5, gpu=True, forkserver=False)(Seq(np.zeros((10, 3)), 32, 48))
    with pytest.raises(AssertionError):
        V(0, gpu=True, forkserver=False)(Seq(ev, 32, 48))
    with pytest.raises(AssertionError):
        V(5, gpu=True, forkserver=False)(Seq(ev, 0, 48))
    with pytest.raises(IndexError):
        V(5, gpu=True, forkserver=False)(Seq(np.zeros((0, 4)), 32, 48))
    bad = ev.copy()
    bad[10, 2] = 4000.0                                           # flat index outside the grid
    with pytest.raises(IndexError):
        V(5, gpu=True, forkserver=False, strict=True)(Seq(bad, 32, 48))
    V(5, gpu=True, forkserver=False)(Seq(bad, 32, 48))            # non-strict: vote dropped, no error


@pytest.mark.parametrize("path", ["pair", "direct", "cluster", "interleaved", "tiled"])
def test_forced_vote_path_parity(golden, V, monkeypatch, path):
    """("tiled": the tile-binned exact kernel forced for the order-free mode as well.)
    The atomic mode has three implementations chosen by size: cluster-resident (a window's grid lives in the
    shared memory of an 8-CTA cluster; MVSEC-sized windows), direct L2 atomics, and the pair layout (>= 2 events
    per voxel).  Force each one and repeat the parity checks, including the reference's row-wrap and
    dropped-vote corner cases (a forced cluster path falls back to direct for grids that do not fit)."""
    monkeypatch.setenv("EEM_VOXEL_PATH", path)
    g = golden("voxel")
    for name in cases_of(g):
        nb, h, w = (int(v) for v in g[f"{name}__shape"])
        ev = g[f"{name}__events"]
        raw = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev.copy(), h, w)).cpu().numpy()
        assert rel_close(raw, g[f"{name}__raw"]).all(), (name, np.abs(raw - g[f"{name}__raw"]).max())
        norm = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(ev.copy(), h, w)).cpu().numpy()
        assert rel_close(norm, g[f"{name}__norm"]).all(), (name, np.abs(norm - g[f"{name}__norm"]).max())
    rng = np.random.default_rng(21)
    for n, nb, h, w, clustered in [(600_000, 15, 720, 1280, True), (200_000, 5, 260, 346, False)]:
        ev = make_events(rng, n, h, w, clustered)
        ref_raw, _, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)
        out = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
        assert rel_close(out, ref_raw).all(), np.abs(out - ref_raw).max()
        ref_norm = ref_ops.voxelize(ev, nb, h, w, normalize=True).numpy()
        out_n = V(nb, gpu=True, normalize=True, forkserver=False)(Seq(ev, h, w)).cpu().numpy()
        assert rel_close(out_n, ref_norm).all(), np.abs(out_n - ref_norm).max()
    # ragged batch through the forced path (more windows than co-resident clusters for the cluster path)
    seqs = [Seq(make_events(rng, n, 64, 96), 64, 96) for n in (1, 2, 777, 5000, 31, 12345) * 7]
    batch = V(5, gpu=True, normalize=True, forkserver=False).voxelize_batch(seqs).cpu().numpy()
    for k, s in enumerate(seqs):
        assert rel_close(batch[k], ref_ops.voxelize(s.features, 5, 64, 96, normalize=True).numpy()).all(), k
    # strict mode still sees dropped votes
    bad = make_events(rng, 2000, 64, 96)
    bad[7, 2] = 5000.0
    with pytest.raises(IndexError):
        V(5, gpu=True, forkserver=False, strict=True)(Seq(bad, 64, 96))
    # out-of-grid votes are dropped (and counted in strict mode) on this path too
    bad = make_events(rng, 1000, 32, 48)
    bad[10, 2] = 4000.0
    bad[20, 1] = -7.0
    bad[20, 2] = 0.0
    with pytest.raises(IndexError):
        V(5, gpu=True, forkserver=False, strict=True)(Seq(bad, 32, 48))
    V(5, gpu=True, forkserver=False)(Seq(bad, 32, 48))


@pytest.mark.parametrize("path,orderfree", [("radix", "0"), ("tiled", "1")])
def test_forced_deterministic_and_orderfree_variants(golden, V, monkeypatch, path, orderfree):
    """The deterministic mode's fallback (stable radix sort, bit-exact) and the tile path's shared-memory-atomic
    variant (order-free, <= 1e-5 relative), forced through the environment."""
    monkeypatch.setenv("EEM_VOXEL_PATH", path)
    monkeypatch.setenv("EEM_VOXEL_ORDERFREE", orderfree)
    g = golden("voxel")
    rng = np.random.default_rng(33)
    extra = [(make_events(rng, 300_000, 260, 346, True), 5, 260, 346), (make_events(rng, 400_000, 720, 1280, False), 15, 720, 1280)]
    cases = [(g[f"{n}__events"], *(int(v) for v in g[f"{n}__shape"])) for n in cases_of(g)] + extra
    for ev, nb, h, w in cases:
        ref, _, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)
        if path == "radix":
            out = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev.copy(), h, w)).cpu().numpy()
            assert np.array_equal(out, ref)
        else:
            out = V(nb, gpu=True, normalize=False, forkserver=False)(Seq(ev.copy(), h, w)).cpu().numpy()
            assert rel_close(out, ref).all(), np.abs(out - ref).max()


def test_deterministic_unusual_polarity_and_off_sensor_events(V):
    """Tile-binned exact path, rare branches: polarity values other than +-1 (kept in the side array), p == 0 (-> -1),
    x >= W (the reference's silent row wrap), votes that leave the grid (dropped and counted), unsorted stamps inside a
    window handed straight to the C ABI layer (chunk ranges then overlap)."""
    rng = np.random.default_rng(41)
    nb, h, w = 7, 100, 150
    ev = make_events(rng, 60_000, h, w)
    ev[:, 3] = rng.choice([-1.0, 0.0, 1.0, 0.5, 2.0, -3.25], size=ev.shape[0])
    ev[::97, 1] = w + rng.integers(0, 40, size=ev[::97].shape[0])          # wraps into the next row / bin
    ev[5::1013, 2] = -3.0                                                  # flat index < 0 for the first bin: dropped
    ref, dropped, _ = c_oracle.voxelize(ev, nb, h, w, normalize=False)
    assert dropped > 0
    out = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True)(Seq(ev.copy(), h, w)).cpu().numpy()
    assert np.array_equal(out, ref)
    with pytest.raises(IndexError):
        V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True, strict=True)(Seq(ev.copy(), h, w))
    # unsorted rows straight into ops.voxelize (EventSequence would sort them): still the reference's result,
    # which votes in ROW order with first/last stamps taken from the first/last row
    from eemflow_b200 import ops
    perm = rng.permutation(ev.shape[0])
    shuffled = np.ascontiguousarray(ev[perm])
    ref_s, _, _ = c_oracle.voxelize(shuffled, nb, h, w, normalize=False)
    d_ev = torch.from_numpy(shuffled).cuda()
    off = torch.tensor([0, shuffled.shape[0]], dtype=torch.int64, device="cuda")
    out_s = ops.voxelize(d_ev, off, shuffled.shape[0], nb, h, w, normalize=False, deterministic=True).cpu().numpy()[0]
    assert np.array_equal(out_s, ref_s)


def test_voxelize_concat_equals_the_dt4_loader_chain(V):
    """dt4 windows (loader/MVSEC.py:245-262): four consecutive frames' events joined, EventSequence(x1e6, relative),
    voxelized.  voxelize_concat does the join (consecutive device rows) and the x1e6 on the device: bit-exact in
    deterministic mode, <= 1e-5 in the order-free mode; an unsorted group takes the host path and still matches."""
    from eemflow_b200 import EventSequence
    rng = np.random.default_rng(77)
    h, w, nb = 260, 346, 5
    groups = []
    for g in range(3):
        t0, parts = 1000.0 + 0.2 * g, []
        for k in range(4):                                        # four frames of ~12.5 ms each, absolute seconds
            n = int(rng.integers(20_000, 40_000))
            t = np.sort(rng.uniform(t0 + 0.0125 * k, t0 + 0.0125 * (k + 1), size=n))
            parts.append(np.stack([t, rng.integers(0, w, n).astype(np.float64), rng.integers(0, h, n).astype(np.float64),
                                   2.0 * rng.integers(0, 2, n) - 1.0], axis=1))
        groups.append(parts)
    refs = []
    for parts in groups:
        seq = EventSequence(None, {"height": h, "width": w}, features=np.concatenate(parts), timestamp_multiplier=1e6,
                            convert_to_relative=True)
        refs.append(c_oracle.voxelize(seq.features, nb, h, w, normalize=False)[0])
    det = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_concat(groups, h, w, timestamp_multiplier=1e6)
    atom = V(nb, gpu=True, normalize=False, forkserver=False).voxelize_concat(groups, h, w, timestamp_multiplier=1e6)
    for g in range(3):
        assert np.array_equal(det[g].cpu().numpy(), refs[g]), g
        assert rel_close(atom[g].cpu().numpy(), refs[g]).all(), g
    shuffled = [groups[0][::-1]] + groups[1:]                      # frames in the wrong order: host path (sort), same result
    det_s = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_concat(shuffled, h, w, timestamp_multiplier=1e6)
    if len(np.unique(np.concatenate(groups[0])[:, 0])) == sum(p.shape[0] for p in groups[0]):
        assert np.array_equal(det_s[0].cpu().numpy(), refs[0])


def test_packed_columns_match_reference_loader_chain(V):
    """HREM .npz style columns (x, y, t [int64 ns], p in {0,1}) through eem_voxelize_soa vs the reference chain
    get_compressed_events -> EventSequence(x1e6, relative) -> voxelizer restated on the CPU: bit-exact in
    deterministic mode, <= 1e-5 relative in atomic mode."""
    from eemflow_b200 import EventSequence
    rng = np.random.default_rng(11)
    h, w, nb = 72, 128, 15
    windows, seqs = [], []
    for n in (5000, 1, 20000):
        t_ns = np.sort(rng.integers(1_000_000_000, 1_050_000_000, size=n)).astype(np.int64)
        cols = {"x": rng.integers(0, w, size=n).astype(np.uint16), "y": rng.integers(0, h, size=n).astype(np.uint16),
                "t": t_ns, "p": rng.integers(0, 2, size=n).astype(np.uint8)}
        windows.append(cols)
        # loader/loader_utils.py:26-37 + EventSequence(timestamp_multiplier=1e6, convert_to_relative=True)
        feats = np.stack([cols["t"] * 1e-9, cols["x"], cols["y"], 2 * cols["p"].astype(np.int64) - 1], axis=1).astype(np.float64)
        seqs.append(EventSequence(None, {"height": h, "width": w}, features=feats, timestamp_multiplier=1e6, convert_to_relative=True))
    det = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_columns(windows, h, w).cpu().numpy()
    atom = V(nb, gpu=True, normalize=True, forkserver=False).voxelize_columns(windows, h, w).cpu().numpy()
    for k, s in enumerate(seqs):
        ref_raw, _, _ = c_oracle.voxelize(s.features, nb, h, w, normalize=False)
        assert np.array_equal(det[k], ref_raw), k
        ref_norm = ref_ops.voxelize(s.features, nb, h, w, normalize=True).numpy()
        assert rel_close(atom[k], ref_norm).all(), k
    # page-locked columns in the device layout (a loader with pin_memory=True) take the staging-free upload: same bits
    def pin(a, dtype):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    pinned = [{"t": pin(c["t"], torch.int64), "x": pin(c["x"], torch.int16), "y": pin(c["y"], torch.int16),
               "p": pin(c["p"], torch.int8)} for c in windows]
    from eemflow_b200.event_utils import _ColumnStage
    assert all(_ColumnStage._pinned_exact(tuple(c[k] for k in ("t", "x", "y", "p")), np.int64) for c in pinned)
    det_p = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_columns(pinned, h, w).cpu().numpy()
    assert np.array_equal(det_p, det)
    pinned[2] = {k: pin(v[::-1].copy(), torch.from_numpy(v).dtype) for k, v in pinned[2].items()}      # unsorted pinned window
    atom_p = V(nb, gpu=True, normalize=True, forkserver=False).voxelize_columns(pinned, h, w).cpu().numpy()
    assert rel_close(atom_p[2], atom[2]).all() and rel_close(atom_p[0], atom[0]).all()
    # unsorted columns are sorted by time stamp first, as EventSequence does (loader/loader_utils.py:365-366)
    perm = rng.permutation(windows[2]["t"].shape[0])
    shuffled = [windows[0], windows[1], {k: v[perm] for k, v in windows[2].items()}]
    if len(np.unique(windows[2]["t"])) == windows[2]["t"].shape[0]:       # argsort is then a unique permutation
        det_s = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_columns(shuffled, h, w).cpu().numpy()
        assert np.array_equal(det_s, det)
    atom_s = V(nb, gpu=True, normalize=True, forkserver=False).voxelize_columns(shuffled, h, w).cpu().numpy()
    assert rel_close(atom_s[2], ref_ops.voxelize(seqs[2].features, nb, h, w, normalize=True).numpy()).all()
    # float64 columns = the values of features[:,0]
    wf = [{"x": s.features[:, 1], "y": s.features[:, 2], "t": s.features[:, 0], "p": s.features[:, 3]} for s in seqs]
    det2 = V(nb, gpu=True, normalize=False, forkserver=False, deterministic=True).voxelize_columns(wf, h, w).cpu().numpy()
    assert np.array_equal(det2, det)


def test_direct_ctypes_binding_as_documented():
    """The raw C-ABI call sequence INTEGRATION.md section 4 shows (no eemflow_b200 Python layer involved)."""
    import ctypes as C
    from pathlib import Path
    lib = C.CDLL(str(Path(__file__).resolve().parent.parent / "eemflow_b200" / "libeemflow_b200.so"))
    lib.eem_voxelize.restype = C.c_int
    lib.eem_voxelize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.eem_voxelize_workspace_bytes.restype = C.c_size_t
    lib.eem_voxelize_workspace_bytes.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.eem_last_error_string.restype = C.c_char_p
    rng = np.random.default_rng(2)
    nb, h, w = 5, 40, 60
    feats = make_events(rng, 4000, h, w)
    ev = torch.from_numpy(feats).cuda()
    off = torch.tensor([0, ev.shape[0]], dtype=torch.int64, device="cuda")
    grid = torch.empty(nb, h, w, device="cuda")
    mode, normalize = 0, 1
    nbytes = lib.eem_voxelize_workspace_bytes(ev.shape[0], 1, nb, h, w, mode, normalize)
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cuda")
    rc = lib.eem_voxelize(ev.data_ptr(), off.data_ptr(), 1, ev.shape[0], ev.shape[0], nb, h, w, mode, normalize,
                          grid.data_ptr(), None, None, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.eem_last_error_string().decode()
    ref = ref_ops.voxelize(feats, nb, h, w, normalize=True).numpy()
    assert rel_close(grid.cpu().numpy(), ref).all()
    # a NULL events pointer is reported through the status code and the error string, never thrown
    rc = lib.eem_voxelize(None, off.data_ptr(), 1, 10, 10, nb, h, w, mode, normalize, grid.data_ptr(), None, None,
                          ws.data_ptr(), nbytes, None)
    assert rc == -1 and b"NULL" in lib.eem_last_error_string()
