"""Voxelization timing at the BASELINE shapes: tile-binned path (default) vs the round-1 paths (EEM_VOXEL_PATH).
CUDA events around `reps` back-to-back calls (median); K1 = un-normalised call, algorithmic bytes 32 N + 4 nb H W."""
import os
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from eemflow_b200 import ops  # noqa: E402

HBM = 6554.6


def make_events(rng, n, h, w, clustered=False):
    gaps = rng.standard_exponential(n + 1)
    t = np.cumsum(gaps[:n])
    t *= 0.05e6 / (t[-1] + gaps[n])
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
    return np.stack([t, x.astype(np.float64), y.astype(np.float64), 2.0 * rng.integers(0, 2, size=n) - 1.0], axis=1)


def timeit(fn, iters=7, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts) * 1e3


def main():
    dev = torch.device("cuda")
    rng = np.random.default_rng(0)
    cases = [("HREM dt1 x4 uniform", 4, 10_000_000, 15, 720, 1280, False), ("HREM dt1 x4 clustered", 4, 10_000_000, 15, 720, 1280, True),
             ("HREM dt1 x1 uniform", 1, 10_000_000, 15, 720, 1280, False),
             ("HREM dt4 x2 uniform", 2, 40_000_000, 15, 720, 1280, False), ("MVSEC dt1 x64", 64, 30_000, 5, 260, 346, False)]
    which = os.environ.get("BENCH_VOXEL_CASES")
    for name, nw, n, nb, h, w, cl in cases:
        if which and not any(k in name for k in which.split(",")):
            continue
        ev = torch.cat([torch.from_numpy(make_events(rng, n, h, w, cl)) for _ in range(nw)]).to(dev)
        off = torch.arange(nw + 1, dtype=torch.int64, device=dev) * n
        algo = nw * (32 * n + 4 * nb * h * w)
        out = torch.empty((nw, nb, h, w), device=dev)
        for label, det, norm in (("K1 order-free", False, False), ("K1 deterministic", True, False), ("K1+K2 order-free", False, True)):
            t = timeit(lambda: ops.voxelize(ev, off, n, nb, h, w, normalize=norm, deterministic=det, out=out))
            print(f"{name:24s} {label:18s} path={os.environ.get('EEM_VOXEL_PATH', 'tiled'):6s} {t:9.1f} us  "
                  f"{algo / t / 1e3:8.1f} GB/s ({algo / t / 1e3 / HBM:5.1%} of HBM peak, K1 bytes)", flush=True)
        del ev, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
