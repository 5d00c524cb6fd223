import numpy as np, torch, sys
sys.path.insert(0, '.')
import eemflow_b200 as E
from eemflow_b200 import ops
rng = np.random.default_rng(0)
# voxel: tile-binned exact path incl. rare branches, batch of ragged windows
h, w, nb = 100, 150, 7
def ev(n):
    t = np.sort(rng.uniform(0, 0.05, n)) * 1e6
    a = np.stack([t - t[0], rng.integers(0, w, n), rng.integers(0, h, n), rng.choice([-1., 0., 1., 0.5], n)], 1).astype(np.float64)
    a[::97, 1] = w + 5; a[5::211, 2] = -3
    return a
class S:
    def __init__(s, f): s.features, s.image_height, s.image_width = f, h, w
seqs = [S(ev(n)) for n in (1, 2, 777, 9000, 31)]
for det in (True, False):
    enc = E.EventSequenceToVoxelGrid_Pytorch(nb, gpu=True, normalize=True, forkserver=False, deterministic=det)
    g = enc.voxelize_batch(seqs); torch.cuda.synchronize()
# packed pyramid + lookup + unpack, odd level sizes
g_ = torch.Generator().manual_seed(0)
f1 = torch.randn(2, 64, 17, 20, generator=g_).cuda(); f2 = torch.randn(2, 64, 17, 20, generator=g_).cuda()
coords = (torch.stack(torch.meshgrid(torch.arange(17), torch.arange(20), indexing='ij')[::-1], 0).float()[None].repeat(2,1,1,1) + 3*torch.randn(2,2,17,20, generator=g_)).cuda()
for r in (4, 2):
    blk = E.CorrBlock(f1, f2, num_levels=4, radius=r, precision="tf32_f16"); out = blk(coords); _ = blk.corr_pyramid
blk = E.CorrBlock(f1, f2, num_levels=4, radius=4, precision="tf32"); out = blk(coords)
# warp-specialised persistent lookup with more batches than 3 x SMs: the stage ring wraps (mbarrier phases flip)
if torch.cuda.get_device_properties(0).multi_processor_count * 4 <= 12 * 50:
    g1 = torch.randn(12, 32, 36, 44, generator=g_).cuda(); g2 = torch.randn(12, 32, 36, 44, generator=g_).cuda()
    cc = (torch.stack(torch.meshgrid(torch.arange(36), torch.arange(44), indexing='ij')[::-1], 0).float()[None].repeat(12,1,1,1) + 3*torch.randn(12,2,36,44, generator=g_)).cuda()
    blk = E.CorrBlock(g1, g2, num_levels=4, radius=4, precision="tf32_f16"); out = blk(cc); out = blk(cc + 0.5)
# local correlation: whole-map kernel, tcgen05 banded GEMM (ragged sizes)
for (B, C, H, W) in ((2, 24, 5, 6), (1, 40, 9, 12), (2, 40, 21, 36)):
    a_, b_ = torch.randn(B, C, H, W, generator=g_).cuda(), torch.randn(B, C, H, W, generator=g_).cuda()
    ops.local_corr(a_, b_, scale=1.0 / C, precision="fp32")
    ops.local_corr(a_, b_, index=list(range(0, 81, 2)), scale=1.0 / C, precision="tf32")
# backward GEMM + multi resize
a = f1.clone().requires_grad_(True); b = f2.clone().requires_grad_(True)
E.CorrBlock(a, b, num_levels=3, radius=4, precision="fp32")(coords).sum().backward()
flows = [torch.randn(2, 2, hh, ww, device='cuda') for hh, ww in ((5, 6), (10, 12), (20, 24))]
E.upsample2d_flows_as(flows, torch.zeros(2, 1, 50, 70, device='cuda'), if_rate=True)
torch.cuda.synchronize(); print("sanitizer workload done")
