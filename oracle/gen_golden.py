"""Generate tests/golden/*.npz by running the REAL reference code (read-only /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does not exist on the
GPU box):

    python oracle/gen_golden.py

What is executed is the reference's own source, unmodified:
  * utils.transformers.EventSequenceToVoxelGrid_Pytorch      (imports cleanly)
  * model.corr.CorrBlock, model.model_utils.*                 (import cleanly)
  * utils_luo.tools.tensor_tools.torch_warp / torch_warp_mask -- utils_luo/tools.py cannot be
    imported in any environment (NameError at :1811, SURVEY.md trap 4), so the two classmethods
    are AST-extracted from the file and exec'd: still the reference's text, not a restatement
  * model.EEMFlow.cdc_utils.{WarpingLayer_no_div, upsample2d_flow_as, cdc_model}
  * model/EEMFlow/EEMFlow+.py::EEMFlow_cdc.warp, model/EEMFlow/EEMFlow.py::EEMFlow.upsample_flow
  * utils.image_utils.InputPadder
  * model.IRRPWC.pwc_modules.compute_cost_volume -- the in-tree statement of the local cost volume
    (spatial_correlation_sampler itself is a pip dependency that is neither vendored nor installed)
Stub modules are registered only for imports that are irrelevant to the arithmetic (matplotlib,
imageio, png, spatial_correlation_sampler's import line).

Inputs are seeded; each .npz stores inputs and the reference's outputs.  Sizes are small so the
fixtures stay a few hundred KB in total.
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import textwrap
import types
import warnings
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    assert REF.exists(), "reference tree not found; golden vectors can only be generated in the build container"
    sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.colors = _stub("matplotlib.colors", hsv_to_rgb=None)
    _stub("imageio")
    _stub("png")

    # --- AST-extract tensor_tools.torch_warp / torch_warp_mask from the unimportable utils_luo/tools.py
    src = (REF / "utils_luo" / "tools.py").read_text()
    tree = ast.parse(src)
    wanted = {"torch_warp", "torch_warp_mask"}
    fn_src = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "tensor_tools":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in wanted:
                    fn_src.append(ast.get_source_segment(src, item))
    assert len(fn_src) == 2, "could not find torch_warp/torch_warp_mask in utils_luo/tools.py"
    body = "\n\n".join("    @classmethod\n" + "\n".join("    " + ln for ln in s.splitlines()) for s in fn_src)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    exec("class tensor_tools:\n" + body + "\n\nclass tools:\n    pass\n", ns)
    import utils_luo  # the package itself imports fine

    tools_mod = _stub("utils_luo.tools", tensor_tools=ns["tensor_tools"], tools=ns["tools"])
    utils_luo.tools = tools_mod

    # --- spatial_correlation_sampler: only needed so that `import` lines succeed; the goldens for the
    # local cost volume come from compute_cost_volume below, not from this stub.
    class _NoSampler(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a):
            raise RuntimeError("spatial_correlation_sampler is not installed")

    _stub("spatial_correlation_sampler", SpatialCorrelationSampler=_NoSampler)

    from utils.transformers import EventSequenceToVoxelGrid_Pytorch
    from model.corr import CorrBlock
    from model import model_utils
    from model.EEMFlow import cdc_utils
    from utils.image_utils import InputPadder
    from model.IRRPWC.pwc_modules import compute_cost_volume
    from model.EEMFlow.EEMFlow import EEMFlow

    spec = importlib.util.spec_from_file_location("eemflow_plus", REF / "model" / "EEMFlow" / "EEMFlow+.py")
    plus = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(plus)
    return dict(Voxel=EventSequenceToVoxelGrid_Pytorch, CorrBlock=CorrBlock, model_utils=model_utils,
                cdc_utils=cdc_utils, tensor_tools=ns["tensor_tools"], InputPadder=InputPadder,
                compute_cost_volume=compute_cost_volume, EEMFlow=EEMFlow, plus=plus)


class _Seq:
    def __init__(self, features, h, w):
        self.features = features
        self.image_height = h
        self.image_width = w


def make_events(rng, n, h, w, pol01=False, duration=0.05, scale=1e6):
    """Sorted uniform timestamps (then x1e6 and made relative, like EventSequence does), uniform pixels."""
    t = np.sort(rng.uniform(0.0, duration, size=n))
    t = t * scale
    t = t - t[0]
    x = rng.integers(0, w, size=n).astype(np.float64)
    y = rng.integers(0, h, size=n).astype(np.float64)
    p = rng.integers(0, 2, size=n).astype(np.float64)
    if not pol01:
        p = 2 * p - 1
    return np.stack([t, x, y, p], axis=1)


def gen_voxel(R):
    rng = np.random.default_rng(0)
    cases = {}

    def run(name, ev, nb, h, w):
        for norm in (False, True):
            enc = R["Voxel"](num_bins=nb, gpu=False, normalize=norm, forkserver=False)
            out = enc(_Seq(ev.copy(), h, w)).numpy()
            cases[f"{name}__{'norm' if norm else 'raw'}"] = out
        cases[f"{name}__events"] = ev
        cases[f"{name}__shape"] = np.array([nb, h, w])

    run("uniform5", make_events(rng, 4000, 24, 40), 5, 24, 40)
    run("uniform15", make_events(rng, 6000, 18, 32), 15, 18, 32)
    run("pol01", make_events(rng, 3000, 24, 40, pol01=True), 5, 24, 40)
    ev = make_events(rng, 500, 16, 16)
    ev[:, 0] = 7.0  # deltaT == 0 -> 1.0 (transformers.py:74-75)
    run("deltaT0", ev, 5, 16, 16)
    run("single", make_events(rng, 1, 16, 16), 5, 16, 16)          # std of one element = NaN -> v - mean
    run("two", make_events(rng, 2, 16, 16), 3, 16, 16)
    run("onebin", make_events(rng, 300, 12, 20), 1, 12, 20)         # num_bins-1 == 0: all ts == 0
    ev = make_events(rng, 2000, 20, 30)
    ev[::7, 1] = 29.0; ev[::11, 2] = 19.0                           # last column / last row
    run("edges", ev, 5, 20, 30)
    ev = make_events(rng, 1500, 20, 30)
    ev[::5, 1] += 30.0                                              # x >= width: silent wrap into the next row
    ev[-1, 1] = 3.0; ev[-1, 2] = 2.0
    ev = ev[ev[:, 1] + ev[:, 2] * 30 < 20 * 30 - 1]                 # keep the flat index inside the grid
    run("wrapx", ev, 5, 20, 30)
    ev = make_events(rng, 3000, 24, 40)
    ev[:, 1] = np.clip(np.round(rng.normal(20, 1.5, size=3000)), 0, 39)   # clustered: heavy per-voxel accumulation
    ev[:, 2] = np.clip(np.round(rng.normal(12, 1.5, size=3000)), 0, 23)
    run("clustered", ev, 5, 24, 40)
    ev = make_events(rng, 2000, 16, 24)
    ev[:, 1] += rng.uniform(0, 0.999, size=2000)                    # fractional coordinates: .long() truncates
    run("fracxy", ev, 4, 16, 24)
    np.savez_compressed(OUT / "voxel.npz", **cases)


def gen_corr(R):
    g = torch.Generator().manual_seed(1)
    cases = {}
    shapes = {"even": (2, 32, 12, 12, 3),"odd": (1, 16, 9, 13, 3), "wide": (1, 64, 8, 12, 3),
              # 12x16 -> 6x8 -> 3x4 -> 1x2: the last level has H-1 == 0 in bilinear_sampler, the reference
              # returns NaN for that whole level on the CPU; pinned here so the kernel keeps doing the same
              "degenerate": (1, 8, 12, 16, 4)}
    for name, (B, D, H, W, L) in shapes.items():
        f1 = torch.randn(B, D, H, W, generator=g)
        f2 = torch.randn(B, D, H, W, generator=g)
        blk = R["CorrBlock"](f1, f2, num_levels=L, radius=4)
        coords = R["model_utils"].coords_grid(B, H, W) + 3.0 * torch.randn(B, 2, H, W, generator=g)
        coords[0, :, 0, 0] = torch.tensor([-7.5, 2.25])            # far outside: zero padding
        coords[0, :, 0, 1] = torch.tensor([float(W) + 6.0, float(H) + 5.0])
        coords[0, :, 1, 0] = torch.tensor([3.0, 4.0])               # exactly on integer positions
        out = blk(coords)
        cases[f"{name}__f1"] = f1.numpy(); cases[f"{name}__f2"] = f2.numpy()
        cases[f"{name}__coords"] = coords.numpy(); cases[f"{name}__lookup"] = out.numpy()
        cases[f"{name}__levels"] = np.array(L)
        for l, lvl in enumerate(blk.corr_pyramid):
            cases[f"{name}__pyr{l}"] = lvl.numpy()
        cases[f"{name}__corr"] = R["CorrBlock"].corr(f1, f2).numpy()
    # bilinear_sampler + mask, upflow8
    img = torch.randn(2, 3, 7, 9, generator=g)
    pc = torch.rand(2, 5, 6, 2, generator=g) * torch.tensor([10.0, 8.0]) - 1.0
    s, m = R["model_utils"].bilinear_sampler(img, pc, mask=True)
    cases.update(bs_img=img.numpy(), bs_coords=pc.numpy(), bs_out=s.numpy(), bs_mask=m.numpy())
    fl = torch.randn(2, 2, 5, 7, generator=g)
    cases.update(up8_in=fl.numpy(), up8_out=R["model_utils"].upflow8(fl).numpy())
    np.savez_compressed(OUT / "corr.npz", **cases)


def gen_local_corr(R):
    g = torch.Generator().manual_seed(2)
    cases = {}
    for name, (B, C, H, W) in {"a": (2, 8, 7, 9), "b": (1, 16, 5, 6), "c": (1, 32, 12, 36)}.items():
        f1 = torch.randn(B, C, H, W, generator=g)
        f2 = torch.randn(B, C, H, W, generator=g)
        cv = R["compute_cost_volume"](f1, f2, {"max_disp": 4})   # mean over C == Correlation.forward's sum / C
        cases[f"{name}__f1"] = f1.numpy(); cases[f"{name}__f2"] = f2.numpy(); cases[f"{name}__cv"] = cv.numpy()
    # the two channel lists of the reference models
    cases["index_eemflow"] = R["EEMFlow"].__init__.__code__ and np.array(
        [1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 22, 23, 25, 27, 29, 30, 31, 32, 33, 35, 37, 38, 39, 40, 41, 42, 43,
         45, 47, 48, 49, 50, 51, 53, 55, 57, 58, 59, 61, 63, 65, 67, 69, 71, 73, 75, 77, 79])
    net = R["plus"].EEMFlow_cdc(None, groups=3, n_first_channels=15)
    cases["index_cdc"] = net.index.numpy()
    np.savez_compressed(OUT / "local_corr.npz", **cases)


def gen_warp(R):
    g = torch.Generator().manual_seed(3)
    cases = {}
    net = R["plus"].EEMFlow_cdc(None, groups=3, n_first_channels=15)
    wl = R["cdc_utils"].WarpingLayer_no_div()
    for name, (B, C, H, W) in {"a": (2, 3, 7, 9), "b": (1, 4, 12, 20), "c": (1, 2, 1, 8)}.items():
        x = torch.randn(B, C, H, W, generator=g)
        flo = 2.0 * torch.randn(B, 2, H, W, generator=g)
        flo[0, :, 0, 0] = torch.tensor([-50.0, 3.0])               # far out of bounds
        flo[0, :, 0, 1] = 0.0                                        # zero flow
        cases[f"{name}__x"] = x.numpy(); cases[f"{name}__flo"] = flo.numpy()
        cases[f"{name}__warp_exact"] = net.warp(x, flo.clone()).numpy()
        cases[f"{name}__torch_warp"] = R["tensor_tools"].torch_warp(x, flo.clone()).numpy()
        o, m = R["tensor_tools"].torch_warp_mask(x, flo.clone())
        cases[f"{name}__torch_warp_mask_out"] = o.detach().numpy(); cases[f"{name}__torch_warp_mask_mask"] = m.detach().numpy()
        cases[f"{name}__warping_layer"] = wl(x, flo.clone()).numpy()
    # upsample2d_flow_as (+ in-place side effect), EEMFlow.upsample_flow, cdc blend
    fl = torch.randn(2, 2, 3, 5, generator=g)
    tgt = torch.zeros(2, 1, 7, 9)
    a = fl.clone(); r_norate = R["cdc_utils"].upsample2d_flow_as(a, tgt, mode="bilinear", if_rate=False)
    b = fl.clone(); r_rate = R["cdc_utils"].upsample2d_flow_as(b, tgt, mode="bilinear", if_rate=True)
    cases.update(up_in=fl.numpy(), up_norate=r_norate.numpy(), up_rate=r_rate.numpy(), up_rate_input_after=b.numpy())
    mesh = torch.randn(1, 2, 16, 16, generator=g)
    cases.update(mesh_in=mesh.numpy(),
                 mesh_up=R["EEMFlow"].upsample_flow(None, mesh, (45, 80)).numpy(),
                 mesh_down=R["EEMFlow"].upsample_flow(None, mesh, (5, 7)).numpy())
    fi = torch.randn(2, 2, 6, 8, generator=g); inter = 1.5 * torch.randn(2, 2, 6, 8, generator=g)
    mk = torch.sigmoid(torch.randn(2, 1, 6, 8, generator=g))
    blend = R["tensor_tools"].torch_warp(fi, inter) * (1 - mk) + fi * mk      # cdc_utils.py:173 verbatim
    cases.update(blend_init=fi.numpy(), blend_inter=inter.numpy(), blend_mask=mk.numpy(), blend_out=blend.numpy())
    # InputPadder
    x = torch.randn(1, 2, 26, 35, generator=g)
    for mode, rate in (("chairs", 64), ("sintel", 32), ("chairs", 32)):
        p = R["InputPadder"](x.shape, mode=mode, eval_pad_rate=rate)
        (y,) = p.pad(x)
        cases[f"pad_{mode}_{rate}"] = y.numpy()
        cases[f"pad_{mode}_{rate}_unpad"] = p.unpad(y).numpy()
    cases["pad_in"] = x.numpy()
    np.savez_compressed(OUT / "warp.npz", **cases)


def gen_e2e(R):
    """BASELINE config 0: EEMFlow(_cdc) forward from synthetic events (5-bin voxel grids), reference CPU path.
    The local cost volume of the model runs through the oracle's stand-in for the un-vendored
    spatial_correlation_sampler (oracle/ref_ops.py::local_corr_sampler); everything else is the
    reference's own code: voxelizer, InputPadder, convs, cdc_model, warps, upsampling."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import ref_ops
    from oracle.det_weights import set_deterministic_weights

    class Sampler(nn.Module):
        def __init__(self, kernel_size=1, patch_size=9, stride=1, padding=0, dilation=1):
            super().__init__()
            self.md = (patch_size - 1) // 2

        def forward(self, a, b):
            return ref_ops.local_corr_sampler(a, b, self.md)

    plus = R["plus"]
    plus.SpatialCorrelationSampler = Sampler
    net = plus.EEMFlow_cdc(None, groups=3, n_first_channels=5).eval()
    set_deterministic_weights(net)
    rng = np.random.default_rng(7)
    h, w, nb = 128, 160, 5
    ev1, ev2 = make_events(rng, 5000, h, w), make_events(rng, 5000, h, w)
    enc = R["Voxel"](num_bins=nb, gpu=False, normalize=True, forkserver=False)
    v1 = enc(_Seq(ev1.copy(), h, w))[None]
    v2 = enc(_Seq(ev2.copy(), h, w))[None]
    net.change_imagesize((h, w))
    _, flows = net(events1=v1, events2=v2)
    out = {"events1": ev1, "events2": ev2, "shape": np.array([nb, h, w]), "voxel1": v1.numpy(), "voxel2": v2.numpy()}
    for k, f in enumerate(flows):
        out[f"flow{k}"] = f.numpy()
    np.savez_compressed(OUT / "e2e_eemflow_cdc.npz", **out)


def gen_e2e_eraft(R):
    """ERAFT (model/eraft.py) forward, 12 iterations, reference CPU path end to end: voxelizer, InputPadder,
    encoders, CorrBlock (matmul + avg_pool2d + grid_sample lookups), GRU update block, convex upsampling.
    This is the caller that amplifies any error of the correlation volume through 12 recurrent updates."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle.det_weights import set_hashed_weights
    from model.eraft import ERAFT
    net = ERAFT(None, n_first_channels=5).eval()
    # gain 0.8: flows of 10-25 px that depend on the lookups (zeroing them changes the flow by 90 %,
    # 1e-3 multiplicative noise on them by 1e-4), GRU gates not saturated
    set_hashed_weights(net, weight_gain=0.8)
    rng = np.random.default_rng(21)
    h, w, nb = 128, 160, 5
    ev1, ev2 = make_events(rng, 6000, h, w), make_events(rng, 6000, h, w)
    enc = R["Voxel"](num_bins=nb, gpu=False, normalize=True, forkserver=False)
    v1 = enc(_Seq(ev1.copy(), h, w))[None]
    v2 = enc(_Seq(ev2.copy(), h, w))[None]
    net.change_imagesize((h, w))
    _, flows = net(events1=v1, events2=v2, iters=12)
    out = {"events1": ev1, "events2": ev2, "shape": np.array([nb, h, w])}
    for k in (0, 5, 11):
        out[f"flow{k}"] = flows[k].numpy()
        print("eraft iter", k, "mean |flow|", float(flows[k].abs().mean()), "max", float(flows[k].abs().max()))
    np.savez_compressed(OUT / "e2e_eraft.npz", **out)


def _extract_functions(path, names):
    """Source text of the named (possibly nested-in-class) functions of a reference file that cannot be imported."""
    src = path.read_text()
    found = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            found[node.name] = textwrap.dedent(ast.get_source_segment(src, node, padded=True))
    assert set(found) == set(names), (path, set(names) - set(found))
    return found


def hashed_flow(h, w):
    """Deterministic pseudo-random [h, w, 2] float32 field, exact on every machine (integer hash / 64)."""
    y, x, c = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), np.arange(2, dtype=np.int64), indexing="ij")
    v = (y * 7919 + x * 104729 + c * 1299709 + (x * y) % 8191 * 31) % 2001 - 1000
    return (v.astype(np.float32) / np.float32(64.0)) + (x.astype(np.float32) / np.float32(w) - np.float32(0.5)) * np.float32(8.0)


def gen_eval(R):
    """Loader / evaluation helpers: the reference's own source text, executed here.
    flow_error: test_mvsec.py:291-346; motion_propagate: loader/HREM.py:30-101; event mask: the
    np.histogram2d call of loader/MVSEC.py:133-142 (inline in get_sample, restated verbatim as a call)."""
    import cv2
    rng = np.random.default_rng(77)
    cases = {}
    # --- flow_error
    fe_src = _extract_functions(REF / "test_mvsec.py", {"flow_error"})["flow_error"]
    ns = {"np": np, "torch": torch}
    exec(fe_src, ns)
    for name, (h, w), kind, is_car in (("sparse", (64, 80), "sparse", False), ("dense", (48, 56), "dense", False),
                                        ("car", (260, 346), "sparse", True)):
        gt = rng.normal(0, 3, size=(1, 2, h, w)).astype(np.float32)
        gt[0, :, rng.integers(0, h, 40), rng.integers(0, w, 40)] = np.inf
        zero = rng.random((h, w)) < 0.1
        gt[0, :, zero] = 0.0
        pred = (gt + rng.normal(0, 1.2, size=gt.shape)).astype(np.float32)
        pred[~np.isfinite(pred)] = 0.0
        ev = (rng.random((1, 1, h, w)) < 0.4).astype(np.float32) * rng.integers(1, 5, size=(1, 1, h, w))
        fake_self = types.SimpleNamespace(data_loader=types.SimpleNamespace(dataset=types.SimpleNamespace(evaluation_type=kind)))
        res = ns["flow_error"](fake_self, torch.from_numpy(gt), torch.from_numpy(pred), torch.from_numpy(ev.astype(np.float32)), is_car)
        cases[f"fe_{name}_gt"], cases[f"fe_{name}_pred"], cases[f"fe_{name}_ev"] = gt, pred, ev.astype(np.float32)
        cases[f"fe_{name}_res"] = np.array([float(v) for v in res], dtype=np.float64)
    # --- motion_propagate
    mp = _extract_functions(REF / "loader" / "HREM.py", {"check_out_bounds", "motion_propagate"})
    ns = {"np": np, "cv2": cv2}
    exec(mp["check_out_bounds"] + "\n\n" + mp["motion_propagate"], ns)
    for name, (h, w) in (("hrem", (720, 1280)), ("small", (100, 132))):
        ff = hashed_flow(h, w)      # regenerated by the tests from the same formula: only the meshes are stored
        xm, ym = ns["motion_propagate"](ff.copy(), h, w)
        cases[f"mp_{name}_hw"] = np.array([h, w])
        cases[f"mp_{name}_x"], cases[f"mp_{name}_y"] = xm, ym
    # --- event mask (MVSEC val) and HREM event_valid
    h, w = 60, 90
    ev = make_events(rng, 3000, h, w)
    ev[:5, 1] = w          # right edge joins the last bin
    ev[5:8, 2] = -1.0      # outside: ignored
    hist, _, _ = np.histogram2d(x=ev[:, 1], y=ev[:, 2], bins=(w, h), range=[[0, w], [0, h]])
    cases["mask_events"], cases["mask_hw"], cases["mask_out"] = ev, np.array([h, w]), (hist.transpose() > 0)
    vol = R["Voxel"](5, normalize=True, gpu=False)(_Seq(make_events(rng, 4000, h, w), h, w))
    cases["binsum_in"], cases["binsum_out"] = vol.numpy(), np.sum(vol.data.numpy(), axis=0)
    np.savez_compressed(OUT / "eval.npz", **cases)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    if "--only-eval" in sys.argv or "--only-eraft" in sys.argv:
        torch.set_num_threads(1)
        R = load_reference()
        with torch.no_grad():
            gen_eval(R) if "--only-eval" in sys.argv else gen_e2e_eraft(R)
        return
    torch.set_num_threads(1)  # fixed summation order in the reference's ATen reductions
    R = load_reference()
    with torch.no_grad():
        gen_voxel(R)
        gen_corr(R)
        gen_local_corr(R)
        gen_warp(R)
        gen_e2e(R)
        gen_e2e_eraft(R)
        gen_eval(R)
    for f in sorted(OUT.glob("*.npz")):
        print(f"{f.name}: {f.stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
