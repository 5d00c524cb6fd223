"""compute-sanitizer workload for the kernels added in the last session of round 2: the tcgen05 batched GEMM of the
pyramid backward (ragged shapes, both operand layouts, enough tiles per CTA that the stage ring and the two TMEM
accumulators wrap), the bilinear_sampler backward and the differentiable-coordinates CorrBlock call."""
import sys
sys.path.insert(0, '.')
import torch
import eemflow_b200 as E
from eemflow_b200 import ops
g_ = torch.Generator().manual_seed(0)
for (batch, M, N, K, bt) in ((24, 64, 1584, 100, True), (24, 64, 1584, 100, False), (2, 256, 20, 300, False), (3, 32, 130, 8, True)):
    A = torch.randn(batch, M, K, generator=g_).cuda()
    B = torch.randn((batch, N, K) if bt else (batch, K, N), generator=g_).cuda()
    C = torch.empty(batch, M, N, device='cuda')
    assert ops.batched_gemm_tf32_supported(A, B, bt)
    ops.batched_gemm_(C, A, B, b_transposed=bt, alpha=0.5, precision="tf32")
    ops.batched_gemm_(C, A, B, b_transposed=bt, alpha=0.5, accumulate=True, precision="tf32")
    ref = torch.bmm(A, B.transpose(1, 2) if bt else B)
    assert (C - ref).abs().max().item() <= 3e-3 * ref.abs().max().item()
f1 = torch.randn(2, 64, 16, 20, generator=g_).cuda().requires_grad_(True)
f2 = torch.randn(2, 64, 16, 20, generator=g_).cuda().requires_grad_(True)
coords = (torch.stack(torch.meshgrid(torch.arange(16), torch.arange(20), indexing='ij')[::-1], 0).float()[None].repeat(2, 1, 1, 1)
          + 3 * torch.randn(2, 2, 16, 20, generator=g_)).cuda()
E.CorrBlock(f1, f2, num_levels=4, radius=4, precision="tf32")(coords).sum().backward()          # TF32 backward GEMMs
c = coords.clone().requires_grad_(True)
E.CorrBlock(f1, f2, num_levels=3, radius=3, precision="fp32")(c).sum().backward()               # coordinates with grad
img = torch.randn(3, 5, 9, 11, generator=g_).cuda().requires_grad_(True)
pts = (torch.rand(3, 7, 6, 2, generator=g_) * 14 - 2).cuda().requires_grad_(True)
E.bilinear_sampler(img, pts).sum().backward()
torch.cuda.synchronize(); print("sanitizer workload done")
