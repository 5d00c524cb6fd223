"""GPU micro-benchmark of the correlation-pyramid kernel with the EEM_TF32_DEBUG timing knobs."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops  # noqa: E402


def run(B, D, H, W, L, label, iters=10, precision="tf32"):
    f1 = torch.randn(B, D, H, W, device="cuda")
    f2 = torch.randn(B, D, H, W, device="cuda")
    out = ops.corr_pyramid(f1, f2, L, precision=precision)
    for _ in range(3):
        ops.corr_pyramid(f1, f2, L, precision=precision, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        ops.corr_pyramid(f1, f2, L, precision=precision, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    P = H * W
    tot = sum((H >> l) * (W >> l) for l in range(L))
    flops = 2.0 * B * P * tot * D
    byts = 4.0 * B * P * tot
    print(f"{label:34s} {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  out {byts/ms/1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    shape = (32, 256, 36, 44, 4)
    for bn, bk in (("256", "32"), ("128", "64"), ("128", "32")):
        os.environ["EEM_TF32_BN"], os.environ["EEM_TF32_BK"] = bn, bk
        for dbg, label in [(0, "full"), (12, "MMA only"), (7, "resident only"), (2, "no stores")]:
            os.environ["EEM_TF32_DEBUG"] = str(dbg)
            run(*shape, f"BN={bn} BK={bk} {label}")
        os.environ["EEM_TF32_DEBUG"] = "0"
        run(4, 256, 92, 160, 4, f"BN={bn} BK={bk} HREM B=4", iters=3)
