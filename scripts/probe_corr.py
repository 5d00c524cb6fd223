"""Timeline probe of the TF32 GEMM's MMA-issuing thread (CTA 0): where do the cycles of a stage go?"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import _lib, ops  # noqa: E402

f1 = torch.randn(32, 256, 36, 44, device="cuda")
f2 = torch.randn(32, 256, 36, 44, device="cuda")
for bn in ("256", "128"):
    os.environ["EEM_TF32_BN"] = bn
    for dbg, label in [(16, "full"), (16 + 12, "MMA only"), (16 + 7, "resident only"), (16 + 2, "no stores")]:
        os.environ["EEM_TF32_DEBUG"] = str(dbg)
        out = ops.corr_pyramid(f1, f2, 4, precision="tf32")
        torch.cuda.synchronize()
        buf = (C.c_longlong * 8192)()
        lib = C.CDLL(str(_lib.LIB_PATH))
        assert lib.eem_debug_read_probe(buf, 8192) == 0
        a = np.array(buf[:], dtype=np.int64).reshape(-1, 4)
        a = a[(a[:, 0] > 0)][8:200]          # skip the first stages (pipeline fill)
        wait = a[:, 1] - a[:, 0]
        issue = a[:, 2] - a[:, 1]
        commit = a[:, 3] - a[:, 2]
        period = np.diff(a[:, 0])
        print(f"BN={bn} {label:14s} stages={len(a)} wait(full) {np.median(wait):6.0f}  issue MMAs {np.median(issue):6.0f}  commit {np.median(commit):6.0f}"
              f"  stage period median {np.median(period):6.0f} mean {period.mean():7.0f} cycles", flush=True)
