"""BASELINE configs[4]: end-to-end EEMFlow_cdc (and ERAFT) inference on one B200, batch sweep at HREM (720x1280,
15 bins) and MVSEC (260x346, 5 bins) resolution: events already voxelized and resident, forward only, random-init
weights, cuDNN convolutions + this library's hot-path kernels.  Prints a markdown table (frame pairs per second).
The reference's CPU path for the same models can only be run where /root/reference exists; its numbers from the
build container are quoted in profiles/r01/README.md."""
import sys, time, statistics
sys.path.insert(0, ".")
import torch
from eemflow_b200.models import EEMFlow_cdc, ERAFT, GraphedInference

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True

def run(model, nb, h, w, batch, iters=5, graphed=False, **kw):
    v1 = torch.randn(batch, nb, h, w, device=dev)
    v2 = torch.randn(batch, nb, h, w, device=dev)
    model.change_imagesize((h, w))
    if graphed:
        g = GraphedInference(model)
        model = lambda events1, events2, **k: g(events1, events2, **k)
    with torch.no_grad():
        for _ in range(2):
            model(events1=v1, events2=v2, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); model(events1=v1, events2=v2, **kw); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)

print("| model | resolution | batch | ms / forward | frame pairs / s | peak mem GB | one CUDA graph: ms | pairs / s |")
print("|---|---|---:|---:|---:|---:|---:|---:|")
for name, ctor, kw in (("EEMFlow_cdc", lambda nb: EEMFlow_cdc(None, groups=3, n_first_channels=nb), {}),
                       ("ERAFT (12 iters, tf32 volume)", lambda nb: ERAFT(None, n_first_channels=nb), {"iters": 12})):
    for res, nb, h, w, batches in (("MVSEC 260x346", 5, 260, 346, (1, 8, 32, 128)), ("HREM 720x1280", 15, 720, 1280, (1, 2, 4, 8, 16, 32))):
        model = ctor(nb).to(dev).eval()
        for B in batches:
            torch.cuda.reset_peak_memory_stats()
            try:
                ms = run(model, nb, h, w, B, **kw)
            except torch.OutOfMemoryError:
                print(f"| {name} | {res} | {B} | OOM | | |", flush=True)
                torch.cuda.empty_cache()
                break
            mem = torch.cuda.max_memory_allocated() / 1e9
            gms = run(model, nb, h, w, B, graphed=True, **kw) if B <= 8 else float("nan")
            print(f"| {name} | {res} | {B} | {ms:.1f} | {B / ms * 1e3:.1f} | {mem:.1f} | {gms:.2f} | {B / gms * 1e3:.1f} |", flush=True)
        del model
        torch.cuda.empty_cache()
