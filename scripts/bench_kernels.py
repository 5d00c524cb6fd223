"""Per-kernel roofline table at the BASELINE.json shapes (MVSEC 260x346 and HREM 720x1280).

Every row: CUDA-event time of the public op (median of `iters` after warm-up, inputs larger than L2
or L2 flushed between iterations), algorithmic bytes / flops as defined in DESIGN.md section 4, and
the fraction of the measured peak (MEASURED_PEAKS.json).  Output: markdown table on stdout.
"""
import json
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import eemflow_b200 as E  # noqa: E402
from eemflow_b200 import ops  # noqa: E402
from eemflow_b200.correlation import EEMFLOW_CDC_INDEX  # noqa: E402

PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TC = PEAKS["hbm_gbs"], PEAKS["bf16_tflops"]
dev = torch.device("cuda")
_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []


def timeit(fn, iters=10, flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()          # 256 MiB > 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e-3


def row(name, shape, secs, nbytes=None, flops=None, note=""):
    gbs = nbytes / secs / 1e9 if nbytes else None
    tf = flops / secs / 1e12 if flops else None
    rows.append((name, shape, secs * 1e6, gbs, gbs / HBM if gbs else None, tf, tf / TC if tf else None, note))
    print(f"{name:28s} {shape:34s} {secs*1e6:9.1f} us"
          + (f"  {gbs:8.1f} GB/s ({gbs/HBM:5.1%} HBM)" if gbs else "")
          + (f"  {tf:7.1f} TFLOP/s ({tf/TC:5.1%} of bf16 peak)" if tf else "") + ("  " + note if note else ""), flush=True)


def make_events(rng, n, h, w, clustered=False):
    t = np.sort(rng.uniform(0.0, 0.05, size=n)) * 1e6
    t -= t[0]
    if clustered:
        k = int(0.8 * n)
        centres = rng.uniform([0, 0], [w, h], size=(8, 2))
        xy = centres[rng.integers(0, 8, size=k)] + rng.normal(0, 0.035 * min(h, w), size=(k, 2))
        x = np.concatenate([np.clip(np.round(xy[:, 0]), 0, w - 1), rng.integers(0, w, size=n - k)])
        y = np.concatenate([np.clip(np.round(xy[:, 1]), 0, h - 1), rng.integers(0, h, size=n - k)])
        perm = rng.permutation(n)
        x, y = x[perm], y[perm]
    else:
        x, y = rng.integers(0, w, size=n), rng.integers(0, h, size=n)
    return np.stack([t, x.astype(np.float64), y.astype(np.float64), 2.0 * rng.integers(0, 2, size=n) - 1.0], 1)


def bench_voxel():
    rng = np.random.default_rng(0)
    for label, n, nb, h, w, nwin, clustered in [("MVSEC dt1 x64 windows", 30_000, 5, 260, 346, 64, False),
                                                ("HREM dt1 uniform", 10_000_000, 15, 720, 1280, 1, False),
                                                ("HREM dt1 clustered", 10_000_000, 15, 720, 1280, 1, True),
                                                ("HREM dt4 uniform", 40_000_000, 15, 720, 1280, 1, False),
                                                ("HREM dt1 5 bins", 10_000_000, 5, 720, 1280, 1, False)]:
        ev = np.concatenate([make_events(rng, n, h, w, clustered) for _ in range(nwin)], 0)
        d_ev = torch.from_numpy(ev).to(dev)
        d_off = torch.arange(0, nwin + 1, dtype=torch.int64, device=dev) * n
        out = torch.empty(nwin, nb, h, w, device=dev)
        vox_bytes = 32 * n * nwin + 4 * nb * h * w * nwin
        import os
        for path in ("direct", "pair"):
            os.environ["EEM_VOXEL_PATH"] = path
            s = timeit(lambda: ops.voxelize(d_ev, d_off, n, nb, h, w, normalize=False, out=out))
            row(f"K1 voxelize atomic/{path}", label, s, vox_bytes)
            s = timeit(lambda: ops.voxelize(d_ev, d_off, n, nb, h, w, normalize=True, out=out))
            row(f"K1+K2 voxelize+normalize/{path}", label, s, vox_bytes + 12 * nb * h * w * nwin)
        os.environ.pop("EEM_VOXEL_PATH")
        s = timeit(lambda: ops.voxelize(d_ev, d_off, n, nb, h, w, normalize=True, out=out))
        row("K1+K2 voxelize+normalize/auto", label, s, vox_bytes + 12 * nb * h * w * nwin)
        # packed columns (eem_voxelize_soa): t f64, x/y int16, p int8 = 13 B/event
        d_t = d_ev[:, 0].contiguous()
        d_x, d_y, d_p = d_ev[:, 1].to(torch.int16), d_ev[:, 2].to(torch.int16), d_ev[:, 3].to(torch.int8)
        s = timeit(lambda: ops.voxelize_soa(d_t, d_x, d_y, d_p, d_off, n, nb, h, w, normalize=True, out=out))
        row("K1+K2 voxelize_soa+normalize/auto", label, s, 13 * n * nwin + 4 * nb * h * w * nwin + 12 * nb * h * w * nwin, note="13 B/event")
        del d_t, d_x, d_y, d_p
        s2 = timeit(lambda: ops.voxel_normalize_(out))
        row("K2 voxel_normalize", label, s2, 12 * nb * h * w * nwin)
        if n * nwin <= 12_000_000:
            s3 = timeit(lambda: ops.voxelize(d_ev, d_off, n, nb, h, w, normalize=False, deterministic=True, out=out), iters=5)
            row("K1d voxelize (deterministic)", label, s3, vox_bytes, note="bit-exact mode")
        del d_ev, out


def bench_corr():
    for label, B, D, H, W in [("MVSEC 36x44 B=32", 32, 256, 36, 44), ("MVSEC crop 32x32 B=32", 32, 256, 32, 32),
                              ("HREM 92x160 B=4", 4, 256, 92, 160)]:
        f1 = torch.randn(B, D, H, W, device=dev)
        f2 = torch.randn(B, D, H, W, device=dev)
        P = H * W
        tot = sum((H >> l) * (W >> l) for l in range(4))
        pyr = ops.corr_pyramid(f1, f2, 4, precision="tf32")
        s = timeit(lambda: ops.corr_pyramid(f1, f2, 4, precision="tf32", out=pyr), flush=False)
        row("K3 corr_pyramid tf32", label, s, 4 * B * P * tot + 8 * B * D * P, 2.0 * B * P * tot * D, note="bytes = volume written + fmaps read")
        coords = (E.coords_grid(B, H, W) + 3 * torch.randn(B, 2, H, W)).to(dev)
        out = ops.corr_lookup(pyr, coords, 4)
        taps, h, w = 0, H, W
        for _ in range(4):
            taps += min(100, h * w)
            h, w = h // 2, w // 2
        s = timeit(lambda: ops.corr_lookup(pyr, coords, 4, out=out))
        row("K5 corr_lookup r=4 L=4", label, s, B * P * (4 * 324 + 8 + 4 * taps))
        del pyr, out


def bench_eemflow():
    for label, B, levels, full in [("MVSEC pad 320x384 B=32", 32, [(64, 5, 6), (64, 10, 12), (64, 20, 24), (64, 40, 48), (32, 80, 96)], (260, 346)),
                                   ("HREM pad 768x1280 B=4", 4, [(64, 12, 20), (64, 24, 40), (64, 48, 80), (64, 96, 160), (32, 192, 320)], (720, 1280))]:
        for (c, h, w) in levels[-2:]:
            f1 = torch.randn(B, c, h, w, device=dev)
            f2 = torch.randn(B, c, h, w, device=dev)
            flo = 2 * torch.randn(B, 2, h, w, device=dev)
            px = B * h * w
            s = timeit(lambda: E.correlation_select(f1, f2, EEMFLOW_CDC_INDEX))
            row("K6 local_corr81 (53 ch)", f"{label} C={c} {h}x{w}", s, px * (8 * c + 4 * 53), 162.0 * c * px, note="fp32 FFMA, not tensor cores")
            s = timeit(lambda: E.warp(f2, flo))
            row("K7 backwarp (exact)", f"{label} C={c} {h}x{w}", s, px * (8 * c + 8))
            s = timeit(lambda: E.WarpingLayer_no_div()(f2, flo))
            row("K7 backwarp+mask (halfpix)", f"{label} C={c} {h}x{w}", s, px * (8 * c + 8))
        c, h, w = levels[1]
        fl = torch.randn(B, 2, h, w, device=dev)
        tgt = torch.empty(B, 1, *full, device=dev)
        s = timeit(lambda: E.upsample2d_flow_as(fl.clone(), tgt, if_rate=True))
        row("K8 flow upsample (rate)", f"{label} {h}x{w}->{full[0]}x{full[1]}", s, 4 * 2 * B * full[0] * full[1])
        mesh = torch.randn(B, 2, 16, 16, device=dev)
        s = timeit(lambda: E.upsample_flow(mesh, full))
        row("K8 meshflow->dense", f"{label} 16x16->{full[0]}x{full[1]}", s, 4 * 2 * B * full[0] * full[1])
        x = torch.randn(B, 5, *full, device=dev)
        pad = E.InputPadder(x.shape, mode="chairs", eval_pad_rate=64)
        s = timeit(lambda: pad.pad(x))
        ph, pw = full[0] + pad._pad[2] + pad._pad[3], full[1] + pad._pad[0] + pad._pad[1]
        row("K9 replicate pad", f"{label} {full[0]}x{full[1]}->{ph}x{pw}", s, 4 * 5 * B * (full[0] * full[1] + ph * pw))


if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    print(f"peaks: HBM {HBM} GB/s, bf16 {TC} TFLOP/s ({torch.cuda.get_device_name(0)})")
    if only in ("", "voxel"):
        bench_voxel()
    if only in ("", "corr"):
        bench_corr()
    if only in ("", "eemflow"):
        bench_eemflow()
    print("\n| kernel | shape | time us | GB/s | of HBM peak | TFLOP/s | of bf16 peak | note |")
    print("|---|---|---:|---:|---:|---:|---:|---|")
    for n, sh, us, g, gf, t, tf_, note in rows:
        print(f"| {n} | {sh} | {us:.1f} | {'' if g is None else f'{g:.0f}'} | {'' if gf is None else f'{gf:.1%}'} | "
              f"{'' if t is None else f'{t:.1f}'} | {'' if tf_ is None else f'{tf_:.1%}'} | {note} |")
