"""Bring-up check of the 2-SM pyramid GEMM (EEM_TF32_PAIR=1) against the 1-SM tf32 kernel and the fp32 kernel."""
import os, sys
sys.path.insert(0, ".")
import torch
from eemflow_b200 import ops
torch.manual_seed(0)
for (B, D, H, W, L) in [(1, 32, 16, 16, 1), (1, 64, 16, 16, 2), (2, 256, 36, 44, 4), (1, 256, 20, 24, 3)]:
    f1 = torch.randn(B, D, H, W, device="cuda"); f2 = torch.randn(B, D, H, W, device="cuda")
    os.environ.pop("EEM_TF32_PAIR", None)
    ref = ops.corr_pyramid(f1, f2, L, precision="fp32")
    one = ops.corr_pyramid(f1, f2, L, precision="tf32")
    os.environ["EEM_TF32_PAIR"] = "1"
    two = ops.corr_pyramid(f1, f2, L, precision="tf32")
    torch.cuda.synchronize()
    for l in range(L):
        e1 = (one[l] - ref[l]).abs().max().item(); e2 = (two[l] - ref[l]).abs().max().item(); d = (two[l] - one[l]).abs().max().item()
        print(f"B{B} D{D} {H}x{W} level {l}: |1sm-fp32| {e1:.2e}  |2sm-fp32| {e2:.2e}  |2sm-1sm| {d:.2e}  ref max {ref[l].abs().max().item():.2f}", flush=True)
print("done")
