"""Run the HREM-shaped voxelization once per distribution (for ncu captures)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops  # noqa: E402
from scripts.bench_kernels import make_events  # noqa: E402

rng = np.random.default_rng(0)
n, nb, h, w = 10_000_000, 15, 720, 1280
for clustered in (False, True):
    ev = torch.from_numpy(make_events(rng, n, h, w, clustered)).cuda()
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    out = torch.empty(1, nb, h, w, device="cuda")
    for _ in range(3):
        ops.voxelize(ev, off, n, nb, h, w, normalize=True, out=out)
    torch.cuda.synchronize()
