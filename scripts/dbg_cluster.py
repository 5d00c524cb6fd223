import os, sys, traceback
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops
for cl in ("1", "2"):
    os.environ["EEM_TF32_CLUSTER"] = cl
    for shape in [(1, 64, 8, 12, 3), (1, 32, 12, 12, 3), (2, 256, 36, 44, 4), (32, 256, 36, 44, 4)]:
        B, D, H, W, L = shape
        f1 = torch.randn(B, D, H, W, device="cuda"); f2 = torch.randn(B, D, H, W, device="cuda")
        try:
            ref = ops.corr_pyramid(f1, f2, L, precision="fp32")
            out = ops.corr_pyramid(f1, f2, L, precision="tf32")
            torch.cuda.synchronize()
            err = max((a - b).abs().max().item() for a, b in zip(out, ref))
            print(f"cluster={cl} shape={shape} max err {err:.3e}", flush=True)
        except Exception as e:
            print(f"cluster={cl} shape={shape} EXC {type(e).__name__}: {e}", flush=True)
