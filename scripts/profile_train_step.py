"""Where a training step of the drop-in models spends its device time (torch profiler, top kernels)."""
import sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from eemflow_b200.models import EEMFlow_cdc, ERAFT

which = sys.argv[1] if len(sys.argv) > 1 else "eemflow_cdc"
B, h, w = (8, 256, 320) if which == "eemflow_cdc" else (4, 256, 320)
torch.manual_seed(0)
net = (EEMFlow_cdc(None, groups=3, n_first_channels=5) if which == "eemflow_cdc" else ERAFT(None, n_first_channels=5)).cuda().train()
net.change_imagesize((h, w))
opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
v1 = torch.randn(B, 5, h, w, device="cuda"); v2 = torch.randn(B, 5, h, w, device="cuda")
def step():
    _, flows = net(events1=v1, events2=v2)
    loss = sum(f.abs().mean() for f in flows)
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
print(f"{which}: B={B} {h}x{w}: {a.elapsed_time(b):.1f} ms per training step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
