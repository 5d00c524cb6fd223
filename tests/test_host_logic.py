"""CPU tests of the host side: C-ABI surface, reference-shaped containers, argument handling.
No compute call is made (no GPU here); the GPU parity tests live in test_gpu_*.py."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "eemflow_b200.h").read_text()
    return sorted(set(re.findall(r"EEM_API\s+[\w\s\*]+?\b(eem_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from eemflow_b200 import _lib
    from eemflow_b200.build import build
    build(verbose=False)                       # no-op when up to date; nvcc cross-compiles without a GPU
    syms = _declared_symbols()
    assert len(syms) >= 18
    h = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/eemflow_b200.h but not exported"
    # the ctypes table binds exactly the header's surface
    assert sorted(_lib.SIGNATURES) == syms
    lib = _lib.lib()
    assert lib.eem_version() >= 100
    assert isinstance(lib.eem_last_error_string(), bytes)


def test_workspace_queries_are_pure_host_functions():
    from eemflow_b200 import _lib
    lib = _lib.lib()
    assert lib.eem_voxelize_workspace_bytes(30_000, 1, 5, 260, 346, _lib.VOXEL_ATOMIC, 0) == 0       # direct RED path
    pair = lib.eem_voxelize_workspace_bytes(40_000_000, 1, 15, 720, 1280, _lib.VOXEL_ATOMIC, 0)  # pair-layout scratch (dt4)
    assert lib.eem_voxelize_workspace_bytes(10_000_000, 1, 15, 720, 1280, _lib.VOXEL_ATOMIC, 0) == 0
    assert pair >= 2 * 15 * 720 * 1280 * 4
    det = lib.eem_voxelize_workspace_bytes(10_000_000, 1, 15, 720, 1280, _lib.VOXEL_DETERMINISTIC, 0)
    # tile-binned exact path: 8-byte records + 4-byte polarity side array per event, chunk tables, chunk ranges
    assert 12 * 10_000_000 <= det < 16 * 10_000_000
    assert lib.eem_voxel_normalize_workspace_bytes(64, 5 * 260 * 346) > 0
    assert lib.eem_corr_pyramid_workspace_bytes(32, 256, 36, 44, 1) == 0
    ws = lib.eem_corr_pyramid_workspace_bytes(32, 256, 36, 44, 4)
    assert ws >= 32 * 256 * (18 * 22 + 9 * 11 + 4 * 5) * 4


def test_bad_arguments_return_status_not_crash():
    from eemflow_b200 import _lib
    lib = _lib.lib()
    # argument validation happens before any CUDA call
    rc = lib.eem_voxelize(None, None, 0, 0, 0, 5, 10, 10, 0, 0, None, None, None, None, 0, None)
    assert rc == _lib.EEM_ERR_BAD_ARG
    assert b"n_windows" in lib.eem_last_error_string()
    with pytest.raises(AssertionError):
        _lib.check(rc)
    rc = lib.eem_corr_lookup(None, 1, 4, 4, 4, 4, None, None, None)
    assert rc == _lib.EEM_ERR_BAD_ARG
    idx = (ctypes.c_int * 2)(3, 3)
    rc = lib.eem_local_corr(1, 1, 1, 1, 4, 4, 4, idx, 2, 1.0, 1, None)   # dummy non-NULL pointers, repeated index
    assert rc == _lib.EEM_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _lib.check(rc)


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing them elsewhere."""
    import eemflow_b200
    from eemflow_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.corr_pyramid(torch.zeros(1, 32, 4, 4), torch.zeros(1, 32, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        eemflow_b200.warp(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(NotImplementedError):
        eemflow_b200.SpatialCorrelationSampler(3, 9, 1, 0, 1)
    if not torch.cuda.is_available():
        # the evaluation helpers accept host arrays (they upload them) but have no CPU implementation either
        with pytest.raises(RuntimeError, match="no CPU path"):
            eemflow_b200.flow_error(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4), torch.ones(1, 1, 4, 4))
        with pytest.raises(RuntimeError, match="no CPU path"):
            eemflow_b200.motion_propagate(np.zeros((32, 32, 2), np.float32), 32, 32)
        from eemflow_b200.models import ERAFT
        net = ERAFT(None, n_first_channels=5).eval()       # builds on the CPU (state-dict work), cannot run there
        net.change_imagesize((64, 64))
        with pytest.raises(RuntimeError, match="no CPU path"), torch.no_grad():
            net(events1=torch.zeros(1, 5, 64, 64), events2=torch.zeros(1, 5, 64, 64))


def test_new_entry_points_validate_arguments():
    from eemflow_b200 import _lib
    lib = _lib.lib()
    assert lib.eem_flow_error(None, None, None, 1, 4, 4, 4, None, None) == _lib.EEM_ERR_BAD_ARG
    assert lib.eem_flow_error(1, 1, None, 1, 4, 4, 9, 1, None) == _lib.EEM_ERR_BAD_ARG          # max_row > height
    assert lib.eem_motion_propagate(1, 1, 64, 64, 64, 3, 1, None) == _lib.EEM_ERR_BAD_ARG       # mesh too large
    assert lib.eem_corr_lookup_backward(1, 1, 1, 4, 4, 4, 7, None, None) == _lib.EEM_ERR_BAD_ARG
    arr = (_lib._vp * 4)(1, 1, 1, 1)
    assert lib.eem_corr_lookup_backward(1, 1, 1, 8, 8, 4, 7, arr, None) == _lib.EEM_ERR_UNSUPPORTED   # radius 7
    assert lib.eem_voxelize_soa(None, 0, None, None, None, None, 1, 0, 0, 5, 4, 4, 0, 0, None, None, None, None, 0, None) == _lib.EEM_ERR_BAD_ARG


def test_product_code_never_imports_the_oracle():
    for f in (ROOT / "eemflow_b200").rglob("*.py"):
        src = f.read_text()
        assert "oracle" not in re.sub(r"#.*", "", src).replace("oracle of", ""), f"{f} mentions the oracle"
    for f in (ROOT / "eemflow_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in f.read_text()


def test_event_sequence_container():
    from eemflow_b200 import EventSequence
    rng = np.random.default_rng(0)
    f = np.stack([rng.uniform(1.0, 2.0, 50), rng.integers(0, 9, 50), rng.integers(0, 7, 50),
                  rng.integers(0, 2, 50)], axis=1).astype(np.float64)
    s = EventSequence(None, {"height": 7, "width": 9}, features=f.copy(), timestamp_multiplier=1e6,
                      convert_to_relative=True)
    assert s.is_sorted() and s.features[0, 0] == 0.0 and len(s) == 50
    order = np.argsort(f[:, 0])
    assert np.allclose(s.features[:, 0], (f[order, 0] * 1e6) - f[order, 0].min() * 1e6)
    both = s + s
    assert len(both) == 100 and both.image_height == 7
    import pandas
    df = pandas.DataFrame(f, columns=["ts", "x", "y", "p"])
    s2 = EventSequence(df, {"height": 7, "width": 9})
    assert list(s2.feature_names) == ["ts", "x", "y", "p"] and s2.is_sorted()


def test_input_padder_geometry():
    from eemflow_b200 import InputPadder
    p = InputPadder((1, 5, 260, 346), mode='chairs', eval_pad_rate=64)
    assert p._pad == [19, 19, 0, 60]                 # 260x346 -> 320x384
    p = InputPadder((1, 5, 720, 1280), mode='chairs', eval_pad_rate=64)
    assert p._pad == [0, 0, 0, 48]
    p = InputPadder((1, 5, 260, 346), mode='sintel', eval_pad_rate=32)
    assert p._pad == [3, 3, 14, 14]
    x = torch.zeros(1, 1, 288, 352)
    assert tuple(p.unpad(x).shape) == (1, 1, 260, 346)


def test_pyramid_shapes_and_tf32_gate():
    from eemflow_b200 import ops
    assert ops.pyramid_level_shapes(36, 44, 4) == [(36, 44), (18, 22), (9, 11), (4, 5)]
    assert ops.pyramid_level_shapes(92, 160, 4) == [(92, 160), (46, 80), (23, 40), (11, 20)]
    assert ops.tf32_supported(256, 36, 44) and ops.tf32_supported(256, 92, 160)
    assert not ops.tf32_supported(16, 9, 13) and not ops.tf32_supported(512, 36, 44)


def test_chunked_sortedness_scan_equals_global_scan():
    """The packed-column ingest checks EventSequence.is_sorted chunk by chunk on its staging threads."""
    from eemflow_b200.event_utils import _chunk_is_sorted
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(1, 60))
        t = np.sort(rng.random(n))
        if n > 1 and rng.random() < 0.5:
            i = int(rng.integers(0, n - 1))
            t[i], t[i + 1] = t[i + 1] + 1e-3, t[i]
        want = bool(np.all(t[:-1] <= t[1:]))
        for chunk in (1, 2, 3, 7, 64):
            got = all(_chunk_is_sorted(t, lo, min(n, lo + chunk)) for lo in range(0, n, chunk))
            assert got == want, (n, chunk)


def test_packed_pyramid_layout_query():
    """eem_corr_packed_layout (pure host function): 4x4-pixel tiles, every level padded to 32 elements, levels back to back."""
    from eemflow_b200 import ops
    for (H, W, L) in [(36, 44, 4), (92, 160, 4), (9, 13, 3), (8, 2, 2), (1, 1, 4)]:
        row, offs, lens = ops.packed_row_elems(H, W, L)
        h, w, off = H, W, 0
        for l in range(L):
            n = ((w + 3) // 4) * ((h + 3) // 4) * 16 if h * w > 0 else 0
            n = (n + 31) // 32 * 32
            assert offs[l] == off and lens[l] == n, (H, W, l)
            off += n
            h, w = h // 2, w // 2
        assert row == off and row % 32 == 0


def test_center_crop_matches_the_mvsec_cropper():
    """loader/MVSEC.py:51,189-193: transforms.CenterCrop((256, 256)) on [.., 260, 346] tensors."""
    import torch
    from eemflow_b200 import center_crop
    x = torch.arange(2 * 260 * 346, dtype=torch.float32).view(2, 260, 346)
    out = center_crop(x, 256)
    top, left = int(round((260 - 256) / 2.0)), int(round((346 - 256) / 2.0))
    assert tuple(out.shape) == (2, 256, 256) and torch.equal(out, x[:, top:top + 256, left:left + 256])
    assert out.data_ptr() == x[:, top:, left:].data_ptr()          # a view, no copy


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_corr_pyramid_backward_plumbing(monkeypatch, mode):
    """Host logic of CorrPyramidFn.backward -- level loop, zero-padded copies of the 9 x 11 level for the tcgen05 GEMM,
    ONE multi-segment product for d fmap1, pooling fold -- with the kernels replaced by torch stand-ins on the CPU,
    against autograd through the oracle's pyramid.  (The kernels themselves are covered by the GPU tests.)"""
    import torch.nn.functional as F
    from eemflow_b200 import autograd as ag, ops
    from oracle import ref_ops
    calls = {"multi": 0, "tf32": 0, "fp32": 0}

    def mm(A, B, bt):
        return torch.bmm(A, B.transpose(1, 2) if bt else B)

    def gemm(C, A, B, *, b_transposed, alpha=1.0, accumulate=False, precision="fp32"):
        if precision == "tf32":
            assert A.shape[2] % 4 == 0 and B.shape[2] % 4 == 0 and A.shape[1] % 32 == 0   # the tcgen05 kernel's rules
        calls[precision] += 1
        C.copy_(alpha * mm(A, B, b_transposed) + (C if accumulate else 0))
        return C

    def gemm_multi(C, As, Bs, *, b_transposed, alpha=1.0, accumulate=False):
        assert 1 <= len(As) <= 6 and all(a.shape[2] % 4 == 0 and b.shape[2] % 4 == 0 for a, b in zip(As, Bs))
        calls["multi"] += 1
        C.copy_(alpha * sum(mm(a, b, b_transposed) for a, b in zip(As, Bs)) + (C if accumulate else 0))
        return C

    def pool_bwd(gin, gout, accumulate=False):
        h, w = gin.shape[-2:]
        up = torch.zeros_like(gin)
        up[..., : 2 * (h // 2), : 2 * (w // 2)] = 0.25 * gout.repeat_interleave(2, -2).repeat_interleave(2, -1)
        gin.copy_(up + (gin if accumulate else 0))
        return gin

    monkeypatch.setattr(ops, "corr_pyramid", lambda a, b, L, precision="fp32": [t.detach() for t in ref_ops.corr_pyramid(a, b, L)])
    monkeypatch.setattr(ops, "batched_gemm_", gemm)
    monkeypatch.setattr(ops, "batched_gemm_tf32_multi_", gemm_multi)
    monkeypatch.setattr(ops, "avg_pool2x2", lambda x: F.avg_pool2d(x, 2, 2))
    monkeypatch.setattr(ops, "avg_pool2x2_backward_", pool_bwd)
    monkeypatch.setenv("EEMFLOW_B200_CORR_BACKWARD", mode)
    gen = torch.Generator().manual_seed(2)
    f1 = torch.randn(1, 32, 36, 44, generator=gen, requires_grad=True)
    f2 = torch.randn(1, 32, 36, 44, generator=gen, requires_grad=True)
    ref = ref_ops.corr_pyramid(f1, f2, 4)                     # 36x44, 18x22, 9x11 (99 columns: padded), 4x5
    gs = [torch.randn(t.shape, generator=gen) for t in ref]
    sum((t * g).sum() for t, g in zip(ref, gs)).backward()
    a = f1.detach().clone().requires_grad_(True)
    b = f2.detach().clone().requires_grad_(True)
    mine = ag.CorrPyramidFn.apply(a, b, 4, "fp32")
    sum((t * g).sum() for t, g in zip(mine, gs)).backward()
    for x, y in ((a.grad, f1.grad), (b.grad, f2.grad)):
        assert (x - y).abs().max().item() <= 1e-4 * y.abs().max().item()
    if mode == "tf32":
        assert calls == {"multi": 1, "tf32": 4, "fp32": 0}      # d fmap1 in one launch, d fmap2 one per level, no FFMA
    else:
        assert calls == {"multi": 0, "tf32": 0, "fp32": 8}
    # only fmap2 needs a gradient: no d fmap1 product at all
    calls.update(multi=0, tf32=0, fp32=0)
    b2 = f2.detach().clone().requires_grad_(True)
    sum((t * g).sum() for t, g in zip(ag.CorrPyramidFn.apply(f1.detach(), b2, 4, "fp32"), gs)).backward()
    assert (b2.grad - f2.grad).abs().max().item() <= 1e-4 * f2.grad.abs().max().item()
    assert calls["multi"] == 0 and calls["tf32"] + calls["fp32"] == 4


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/eemflow_b200.h against the ctypes table of eemflow_b200/_lib.py: same number of
    arguments, same class per argument (pointer / int / int64 / size_t / float / double) and the same return class --
    a drifted binding would otherwise only show up as garbage arguments on the GPU."""
    from eemflow_b200 import _lib
    text = re.sub(r"/\*.*?\*/", " ", (ROOT / "include" / "eemflow_b200.h").read_text(), flags=re.S)
    protos = re.findall(r"EEM_API\s+([\w\s\*]+?)\b(eem_\w+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(_lib.SIGNATURES)

    def c_class(decl: str) -> str:
        decl = decl.strip()
        if "*" in decl or "eem_stream_t" in decl:
            return "ptr"
        base = re.sub(r"\bconst\b", "", decl).split()
        kind = base[0] if len(base) <= 2 else " ".join(base[:-1])
        return {"int": "int", "int64_t": "i64", "size_t": "size", "float": "float", "double": "double",
                "long long": "i64", "void": "void"}[kind]

    def ct_class(t) -> str:
        if t is None:
            return "void"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type(ctypes.POINTER(ctypes.c_int))):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_int64: "i64", ctypes.c_longlong: "i64", ctypes.c_size_t: "size",
                ctypes.c_float: "float", ctypes.c_double: "double"}[t]

    for ret, name, args in protos:
        res, argtypes = _lib.SIGNATURES[name]
        want = [] if args.strip() in ("", "void") else [c_class(a) for a in args.split(",")]
        got = [ct_class(t) for t in argtypes]
        assert got == want, f"{name}: header {want} vs ctypes {got}"
        assert ct_class(res) == c_class(ret + " x"), f"{name}: return type"
