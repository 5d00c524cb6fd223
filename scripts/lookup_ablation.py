"""Packed lookup: phase ablations (EEM_LOOKUP_DEBUG bits: 1 no stores, 2 no tile copies, 4 no interpolation) at MVSEC B = 32.
Graph of 12 launches on different coords, as in the step."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops  # noqa: E402

B, H, W, D, L = 32, 36, 44, 256, 4
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(B, D, H, W, device=dev, generator=g)
f2 = torch.randn(B, D, H, W, device=dev, generator=g)
base = torch.stack(torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")[::-1], 0).float()
coords = [base[None] + 3.0 * torch.randn(B, 2, H, W, device=dev, generator=g) for _ in range(12)]
packed = ops.corr_pyramid_packed(f1, f2, L)
out = ops.corr_lookup_packed(packed, coords[0], L, 4)
for dbg in os.environ.get("ABLATE", "0 1 2 3 4 6 7").split():
    os.environ["EEM_LOOKUP_DEBUG"] = dbg
    for c in coords:
        ops.corr_lookup_packed(packed, c, L, 4, out=out)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for c in coords:
            ops.corr_lookup_packed(packed, c, L, 4, out=out)
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gr.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 12 * 1e3)
    print(f"EEM_LOOKUP_DEBUG={dbg}: {best:6.1f} us per launch", flush=True)
