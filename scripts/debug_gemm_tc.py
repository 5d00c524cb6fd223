"""Diagnostic for the tcgen05 batched GEMM (csrc/gemm_tc.cu): the cases run in a child process under a timeout
(a wrong barrier protocol hangs instead of failing), prints the error against an fp64 product, decodes the operand
layouts with one-hot / integer-coded inputs (which (row, k) element the tensor core really used for each k), and times
the MVSEC / HREM backward shapes against the exact FFMA kernel.

    python scripts/debug_gemm_tc.py            # all cases
    python scripts/debug_gemm_tc.py all | timing   (the children)
    ncu --set full -k regex:gemm_tf32 -c 2 ... python scripts/debug_gemm_tc.py ncu
"""
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run_case(batch, M, N, K, bt):
    import torch
    from eemflow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(batch * 7 + M + N + K)
    A = torch.randn(batch, M, K, device="cuda", generator=g)
    B = torch.randn((batch, N, K) if bt else (batch, K, N), device="cuda", generator=g)
    assert ops.batched_gemm_tf32_supported(A, B, bool(bt)), "shape not supported by the tcgen05 kernel"
    C = torch.full((batch, M, N), float("nan"), device="cuda")
    ops.batched_gemm_(C, A, B, b_transposed=bool(bt), alpha=0.5, precision="tf32")
    torch.cuda.synchronize()
    ref = 0.5 * torch.bmm(A.double(), B.double().transpose(1, 2) if bt else B.double())
    err = (C.double() - ref).abs()
    scale = ref.abs().max().item()
    rel = err.max().item() / scale
    nan = int(torch.isnan(C).sum().item())
    print(f"  max rel err {rel:.3e}  nan {nan}", end="")
    if rel > 2e-3 or nan:
        # where are the errors? per 128-column tile, per 32-row chunk, per sample
        e = err.nan_to_num(1e9)
        per_b = e.amax(dim=(1, 2)).tolist()
        per_m = e.amax(dim=(0, 2)).reshape(-1, 32).amax(dim=1).tolist()
        per_n = [e[:, :, n0:n0 + 128].amax().item() for n0 in range(0, N, 128)]
        print(f"\n    per sample {['%.2g' % v for v in per_b]}\n    per 32-row chunk {['%.2g' % v for v in per_m]}"
              f"\n    per 128-col tile {['%.2g' % v for v in per_n[:16]]}", end="")
    print()
    # accumulate on top
    C2 = C.clone()
    ops.batched_gemm_(C2, A, B, b_transposed=bool(bt), alpha=0.25, accumulate=True, precision="tf32")
    torch.cuda.synchronize()
    rel2 = (C2.double() - 1.5 * ref).abs().max().item() / scale
    print(f"  accumulate: max rel err {rel2:.3e}")
    return rel <= 2e-3 and rel2 <= 3e-3 and nan == 0


def decode(bt):
    """Which element does the tensor core use at position k of each operand?  One operand is one-hot in k, the other
    carries row * 32 + k (exact in TF32 below 2048)."""
    import torch
    from eemflow_b200 import ops
    M, N, K = 32, 64, 32
    rows_b = torch.arange(N, device="cuda", dtype=torch.float32)[:, None] * 32 + torch.arange(K, device="cuda", dtype=torch.float32)[None]
    rows_a = torch.arange(M, device="cuda", dtype=torch.float32)[:, None] * 32 + torch.arange(K, device="cuda", dtype=torch.float32)[None]
    Bcode = (rows_b if bt else rows_b.t().contiguous())[None]        # API B: value(n, k) = 32 n + k
    Acode = rows_a[None]                                             # API A: value(m, k) = 32 m + k
    bad = 0
    lines = []
    for name in ("B (tcgen05 A operand)", "A (tcgen05 B operand)"):
        ks, rows_wrong = [], 0
        for k in range(K):
            if name[0] == "B":
                A = torch.zeros(1, M, K, device="cuda")
                A[0, :, k] = 1.0
                C = torch.empty(1, M, N, device="cuda")
                ops.batched_gemm_(C, A, Bcode, b_transposed=bool(bt), precision="tf32")
                got = C[0, 0].round().long()                    # row m = 0: value(n, k') for every n
                want_rows = torch.arange(N, device="cuda")
            else:
                B = torch.zeros((1, N, K) if bt else (1, K, N), device="cuda")
                if bt:
                    B[0, :, k] = 1.0
                else:
                    B[0, k, :] = 1.0
                C = torch.empty(1, M, N, device="cuda")
                ops.batched_gemm_(C, Acode, B, b_transposed=bool(bt), precision="tf32")
                got = C[0, :, 0].round().long()                 # column n = 0: value(m, k') for every m
                want_rows = torch.arange(M, device="cuda")
            ks.append(int(got[1].item()) % 32 if got.numel() > 1 else -1)
            rows_wrong += int(((got // 32) != want_rows).sum().item())
            bad += int((got != want_rows * 32 + k).sum().item())
        lines.append(f"  {name}: k used for k = 0..31: {ks}  rows wrong: {rows_wrong}")
    torch.cuda.synchronize()
    print("\n".join(lines))
    return bad == 0


def timing():
    import torch
    from eemflow_b200 import ops
    for tag, batch, D, P, levels in (("MVSEC B=32 36x44", 32, 256, 1584, (1584, 396, 20)), ("HREM B=2 92x160", 2, 256, 14720, (14720, 3680, 920, 220))):
        for prec in ("tf32", "fp32"):
            if prec == "fp32" and P > 2000:
                continue
            f1 = torch.randn(batch, D, P, device="cuda")
            d1 = torch.empty(batch, D, P, device="cuda")
            ms1 = ms2 = 0.0
            for Pl in levels:
                f2l = torch.randn(batch, D, Pl, device="cuda")
                G = torch.randn(batch, P, Pl, device="cuda")
                d2 = torch.empty(batch, D, Pl, device="cuda")
                for which in (1, 2):
                    def call():
                        if which == 1:
                            ops.batched_gemm_(d1, f2l, G, b_transposed=True, alpha=0.0625, precision=prec)
                        else:
                            ops.batched_gemm_(d2, f1, G, b_transposed=False, alpha=0.0625, precision=prec)
                    call()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        call()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 3
                    if which == 1:
                        ms1 += ms
                    else:
                        ms2 += ms
                del f2l, G, d2
            if prec == "tf32":                     # d_fmap1 as ONE launch over all levels
                As = [torch.randn(batch, D, Pl, device="cuda") for Pl in levels]
                Bs = [torch.randn(batch, P, Pl, device="cuda") for Pl in levels]
                ops.batched_gemm_tf32_multi_(d1, As, Bs, b_transposed=True, alpha=0.0625)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    ops.batched_gemm_tf32_multi_(d1, As, Bs, b_transposed=True, alpha=0.0625)
                e1.record()
                torch.cuda.synchronize()
                print(f"  {tag}: d_fmap1, all levels in one launch {e0.elapsed_time(e1) / 3:.3f} ms")
                del As, Bs
            flop = 2.0 * batch * D * P * sum(levels)
            print(f"  {tag} {prec}: d_fmap1 {ms1:.3f} ms ({flop / ms1 / 1e9:.0f} TFLOP/s)  d_fmap2 {ms2:.3f} ms ({flop / ms2 / 1e9:.0f} TFLOP/s)")


CASES = [
    (1, 32, 128, 8, 1), (1, 32, 128, 8, 0),          # one MMA, one tile
    (1, 32, 128, 32, 1), (1, 32, 128, 32, 0),        # four MMAs: the K advance inside a stage
    (1, 256, 128, 64, 1), (1, 256, 128, 64, 0),      # two stages, all TMEM columns
    (2, 64, 300, 200, 1), (2, 128, 300, 200, 0),     # ragged everything, several samples
    (3, 256, 1584, 1584, 1), (3, 256, 1584, 1584, 0),
    (2, 256, 1584, 20, 1), (2, 256, 20, 1584, 0),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "all":        # child: every case in one process, progress flushed line by line
        failed = 0
        for k, c in enumerate(CASES):
            print("case", *c, flush=True)
            try:
                failed += 0 if run_case(*c) else 1
            except Exception as exc:                      # noqa: BLE001 -- a diagnostic: report and go on
                failed += 1
                print("  EXCEPTION", repr(exc)[:300], flush=True)
            if k == 5:
                for bt in (1, 0):
                    print("decode b_transposed =", bt, flush=True)
                    failed += 0 if decode(bt) else 1
        print(f"{failed} failing", flush=True)
        sys.exit(1 if failed else 0)
    if len(sys.argv) > 1 and sys.argv[1] == "timing":
        timing()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":        # one d_fmap1 and one d_fmap2 launch at MVSEC B = 32, level 0
        import torch
        from eemflow_b200 import ops
        f = torch.randn(32, 256, 1584, device="cuda")
        G = torch.randn(32, 1584, 1584, device="cuda")
        d = torch.empty(32, 256, 1584, device="cuda")
        ops.batched_gemm_(d, f, G, b_transposed=True, alpha=0.0625, precision="tf32")
        ops.batched_gemm_(d, f, G, b_transposed=False, alpha=0.0625, precision="tf32")
        torch.cuda.synchronize()
        sys.exit(0)
    t0 = time.time()
    try:
        rc = subprocess.run([sys.executable, __file__, "all"], timeout=120).returncode
    except subprocess.TimeoutExpired:
        rc = 124
        print("TIMEOUT (hang) in the case printed last", flush=True)
    print(f"rc {rc} after {time.time() - t0:.0f} s", flush=True)
    if rc == 0:
        try:
            subprocess.run([sys.executable, __file__, "timing"], timeout=120)
        except subprocess.TimeoutExpired:
            print("timing: TIMEOUT", flush=True)
    sys.exit(rc)
