"""f32 pyramid + lookup vs the fp16 working pyramid (precision="tf32_f16") at the BASELINE shapes.
CUDA-event time of `reps` back-to-back launches (median over `iters`), so launch latency is amortised as in the
step graph; inputs/outputs are larger than L2."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from eemflow_b200 import ops  # noqa: E402


def timeit(fn, iters=10, reps=6):
    for _ in range(3):
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
    for (B, H, W) in [(32, 36, 44), (2, 92, 160)]:
        D, L = 256, 4
        g = torch.Generator(device="cuda").manual_seed(0)
        f1 = torch.randn(B, D, H, W, device=dev, generator=g)
        f2 = torch.randn(B, D, H, W, device=dev, generator=g)
        base = torch.stack(torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")[::-1], 0).float()
        coords = base[None] + 3.0 * torch.randn(B, 2, H, W, device=dev, generator=g)
        pyr = ops.corr_pyramid(f1, f2, L, precision="tf32")
        t_pyr = timeit(lambda: ops.corr_pyramid(f1, f2, L, precision="tf32", out=pyr))
        out = ops.corr_lookup(pyr, coords, 4)
        t_look = timeit(lambda: ops.corr_lookup(pyr, coords, 4, out=out))
        packed = ops.corr_pyramid_packed(f1, f2, L)
        t_ppyr = timeit(lambda: ops.corr_pyramid_packed(f1, f2, L, out=packed))
        out2 = ops.corr_lookup_packed(packed, coords, L, 4)
        t_plook = timeit(lambda: ops.corr_lookup_packed(packed, coords, L, 4, out=out2))
        err = (out2 - out).abs().max().item()
        P = H * W
        flops = 2.0 * B * P * sum((H >> l) * (W >> l) for l in range(L)) * D
        print(f"B={B} {H}x{W}: f32 pyramid {t_pyr:8.1f} us ({flops / t_pyr / 1e6:6.1f} TFLOP/s)  lookup {t_look:7.1f} us | "
              f"fp16 pyramid {t_ppyr:8.1f} us ({flops / t_ppyr / 1e6:6.1f} TFLOP/s)  lookup {t_plook:7.1f} us | "
              f"max |lookup diff| {err:.3e}", flush=True)


if __name__ == "__main__":
    main()
