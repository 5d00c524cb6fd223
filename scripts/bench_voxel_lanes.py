"""Experiment: time-lane count vs vote-kernel time for the direct and pair layouts (HREM shapes)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops  # noqa: E402
from scripts.bench_kernels import make_events, timeit  # noqa: E402

rng = np.random.default_rng(0)
nb, h, w = 15, 720, 1280
for label, n, clustered in [("dt1 uniform", 10_000_000, False), ("dt1 clustered", 10_000_000, True), ("dt4 uniform", 40_000_000, False)]:
    ev = torch.from_numpy(make_events(rng, n, h, w, clustered)).cuda()
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    out = torch.empty(1, nb, h, w, device="cuda")
    for path in ("direct", "pair"):
        os.environ["EEM_VOXEL_PATH"] = path
        res = []
        for lanes in (1, 2, 4, 8, 16, 64):
            os.environ["EEM_VOXEL_LANES"] = str(lanes)
            t = timeit(lambda: ops.voxelize(ev, off, n, nb, h, w, normalize=False, out=out), iters=5)
            res.append(f"{lanes}:{t*1e6:.0f}")
        print(f"{label:14s} {path:6s} raw      lanes:us  " + "  ".join(res), flush=True)
    os.environ.pop("EEM_VOXEL_LANES")
    for path in ("direct", "pair"):
        os.environ["EEM_VOXEL_PATH"] = path
        t = timeit(lambda: ops.voxelize(ev, off, n, nb, h, w, normalize=True, out=out), iters=5)
        print(f"{label:14s} {path:6s} +normalize default lanes: {t*1e6:.0f} us", flush=True)
    del ev, out
