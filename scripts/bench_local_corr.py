"""Local 9x9 correlation: FFMA kernel vs the tcgen05 banded-GEMM kernel on the EEMFlow_cdc pyramid shapes.
CUDA-event timing, buffers rotated through > L2 worth of data between launches."""
import sys
import torch

sys.path.insert(0, ".")
from eemflow_b200 import ops  # noqa: E402
from eemflow_b200.correlation import EEMFLOW_CDC_INDEX  # noqa: E402

import os
SHAPES = {"mvsec": [(32, 64, 5, 6), (32, 64, 10, 12), (32, 64, 20, 24), (32, 64, 40, 48), (32, 32, 80, 96)],
          "hrem": [(2, 64, 12, 20), (2, 64, 24, 40), (2, 64, 48, 80), (2, 64, 96, 160), (2, 32, 192, 320)]}


def time_it(fn, sets, iters=20):
    """One CUDA graph of `iters` launches (no host launch overhead in the number), replayed three times."""
    for s in sets:
        fn(*s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(*sets[i % len(sets)])
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


for name, shapes in SHAPES.items():
    if os.environ.get("BENCH_LC_ONLY") and name != os.environ["BENCH_LC_ONLY"]:
        continue
    tot = {"fp32": 0.0, "tf32": 0.0}
    for (B, C, H, W) in shapes:
        if os.environ.get("BENCH_LC_SHAPE") and os.environ["BENCH_LC_SHAPE"] != f"{H}x{W}":
            continue
        nbytes = B * C * H * W * 4 * 2 + B * 53 * H * W * 4
        nsets = max(2, min(64, int(300e6 // nbytes) + 1))
        sets = [(torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda"),
                 torch.empty(B, 53, H, W, device="cuda")) for _ in range(nsets)]
        line = f"{name} B={B} C={C} {H}x{W}: "
        for prec in ("fp32", "tf32"):
            if prec == "tf32" and not ops.local_corr_tf32_supported(B, C, H, W):
                line += " tf32: n/a (FFMA kernel)"
                tot[prec] += t
                continue
            t = time_it(lambda a, b, o: ops.local_corr(a, b, index=EEMFLOW_CDC_INDEX, scale=1.0 / C, out=o, precision=prec), sets)
            tot[prec] += t
            line += f" {prec}: {t:7.1f} us ({nbytes / t / 1e3:6.0f} GB/s)"
        print(line, flush=True)
    print(f"{name} total: fp32 {tot['fp32']:.1f} us, tf32 {tot['tf32']:.1f} us", flush=True)
