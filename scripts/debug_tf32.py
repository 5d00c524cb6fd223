"""GPU diagnostic for the tcgen05 TF32 correlation kernel: error maps per 32x32 block against fp64."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from eemflow_b200 import ops  # noqa: E402


def check(B, D, H, W, L=1, seed=0, structured=False):
    g = torch.Generator().manual_seed(seed)
    P = H * W
    if structured:   # f1[d,i] = 1 if d == i % D, f2[d,j] = j + 1 when d == 0 ... exposes permutations
        f1 = torch.zeros(B, D, P)
        f2 = torch.zeros(B, D, P)
        for i in range(P):
            f1[:, i % D, i] = 1.0
        f2[:] = (torch.arange(D).float()[:, None] * 1000 + torch.arange(P).float()[None, :])[None]
        f1, f2 = f1.view(B, D, H, W), f2.view(B, D, H, W)
    else:
        f1 = torch.randn(B, D, H, W, generator=g)
        f2 = torch.randn(B, D, H, W, generator=g)
    f1c, f2c = f1.cuda(), f2.cuda()
    ref = ops.corr_pyramid(f1c, f2c, L, precision="fp32")
    out = ops.corr_pyramid(f1c, f2c, L, precision="tf32")
    torch.cuda.synchronize()
    ok = True
    for l, (a, b) in enumerate(zip(out, ref)):
        a2, b2 = a.view(B, P, -1), b.view(B, P, -1)
        err = (a2 - b2).abs()
        sig = b2.std().item() + 1e-9
        mx = err.max().item()
        print(f"  B={B} D={D} {H}x{W} level {l}: max err {mx:.3e}  rms {err.pow(2).mean().sqrt().item():.3e}  sigma {sig:.3e}"
              f"  nan {torch.isnan(a2).sum().item()}")
        if not (mx <= 1e-2 * max(sig, 1.0)):
            ok = False
            e0 = err[0]
            nb_i, nb_j = (e0.shape[0] + 31) // 32, (e0.shape[1] + 31) // 32
            print("   block max-err map (rows: i/32, cols: j/32), first 8x8:")
            for bi in range(min(nb_i, 8)):
                row = [e0[bi * 32:(bi + 1) * 32, bj * 32:(bj + 1) * 32].max().item() for bj in range(min(nb_j, 8))]
                print("    " + " ".join(f"{v:9.2e}" for v in row))
            print("   out[0,:4,:8]:\n", a2[0, :4, :8].cpu())
            print("   ref[0,:4,:8]:\n", b2[0, :4, :8].cpu())
    return ok


if __name__ == "__main__":
    import os
    torch.manual_seed(0)
    if "--sweep" in sys.argv:
        for variant in (0, 3, 1, 4, 2):
            os.environ["EEM_TF32_VARIANT"] = str(variant)
            print(f"===== variant {variant}")
            try:
                ok = check(1, 32, 8, 16) and check(1, 64, 16, 16)
            except Exception as exc:  # keep sweeping
                print("  exception:", exc)
                ok = False
            print(f"===== variant {variant}:", "OK" if ok else "MISMATCH")
        os.environ.pop("EEM_TF32_VARIANT", None)
    allok = True
    for args in [(1, 32, 8, 16), (1, 32, 16, 16), (1, 64, 16, 16), (1, 256, 16, 24), (2, 256, 36, 44), (1, 128, 23, 40)]:
        print("random", args)
        allok &= check(*args, L=1)
    print("structured (1,32,8,16)")
    check(1, 32, 8, 16, structured=True)
    print("pyramid (2,256,36,44) L=4")
    allok &= check(2, 256, 36, 44, L=4)
    print("TF32 RESULT:", "OK" if allok else "MISMATCH")
