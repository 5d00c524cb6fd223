"""Per-op device time of the EEMFlow_cdc op sequence, launch overhead amortised: every op is captured N times on
rotating buffers (so inputs come from HBM, not L2) into one CUDA graph and the replay is event-timed."""
import sys, statistics
sys.path.insert(0, ".")
import torch
import bench
import eemflow_b200 as E
from eemflow_b200.correlation import EEMFLOW_CDC_INDEX as IDX

dev = torch.device("cuda:0")
B = 32
N = 8
g = torch.Generator(device="cuda").manual_seed(0)
R = lambda *s, sc=1.0: [torch.randn(*s, device=dev, generator=g) * sc for _ in range(N)]

def timed(name, fn, nbytes):
    for k in range(N): fn(k)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for k in range(N): fn(k)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / N * 1e3)
    t = statistics.median(ts)
    print(f"{name:44s} {t:8.1f} us  {nbytes/t/1e3:8.1f} GB/s ({nbytes/t/1e3/6554.6:5.1%})", flush=True)

target = torch.empty(B, 1, bench.H, bench.W, device=dev)
wl = E.WarpingLayer_no_div()
for (c, h, w) in bench.EEM_LEVELS:
    f1, f2, p2 = R(B, c, h, w), R(B, c, h, w), R(B, 32, h, w)
    flow, inter = R(B, 2, h, w, sc=2.0), R(B, 2, h, w, sc=1.5)
    mask = [torch.sigmoid(m) for m in R(B, 1, h, w)]
    small = R(B, 2, h // 2, w // 2, sc=2.0)
    px = B * h * w
    timed(f"local_corr53 C={c} {h}x{w}", lambda k: E.correlation_select(f1[k], f2[k], IDX), px * (8 * c + 4 * 53))
    timed(f"warp exact C={c} {h}x{w}", lambda k: E.warp(f2[k], flow[k]), px * (8 * c + 8))
    timed(f"warp+mask halfpix C=32 {h}x{w}", lambda k: wl(p2[k], flow[k]), px * (8 * 32 + 8))
    timed(f"cdc_blend {h}x{w}", lambda k: E.cdc_blend(flow[k], inter[k], mask[k]), px * 28)
    timed(f"upsample x2 (+in-place rate) -> {h}x{w}", lambda k: E.upsample2d_flow_as(small[k], f1[k], if_rate=True), px * 8 + px * 4)
    timed(f"final upsample {h}x{w} -> {bench.H}x{bench.W}", lambda k: E.upsample2d_flow_as(flow[k], target, if_rate=True), B * 2 * bench.H * bench.W * 4 + px * 16)
    timed(f"clone flow {h}x{w} (torch)", lambda k: flow[k].clone(), px * 16)
