"""Where the end-to-end step time goes: link bandwidth, event staging, overlapped step."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench

dev = torch.device("cuda:0")
def tm(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

big = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
dbig = torch.empty_like(big, device=dev)
ms = tm(lambda: dbig.copy_(big, non_blocking=True)); print(f"H2D 256 MiB pinned: {ms:.2f} ms = {big.numel()/ms/1e6:.1f} GB/s")
ms = tm(lambda: big.copy_(dbig, non_blocking=True)); print(f"D2H 256 MiB pinned: {ms:.2f} ms = {big.numel()/ms/1e6:.1f} GB/s")
s2 = torch.cuda.Stream()
def both():
    dbig[:128 << 20].copy_(big[:128 << 20], non_blocking=True)
    with torch.cuda.stream(s2):
        big[128 << 20:].copy_(dbig[128 << 20:], non_blocking=True)
ms = tm(both); print(f"H2D 128 + D2H 128 MiB concurrently: {ms:.2f} ms")

inp = bench.make_host_inputs(32, 12, 0, pin=True)
step = bench.B200Step(inp, dev, 12)
print("h2d bytes", bench.h2d_bytes(inp) / 1e6, "MB; events", sum(e.nbytes for e in inp["events"]) / 1e6)
ms = tm(lambda: step.enc.voxelize_batch(step.seqs)); print(f"voxelize_batch from host rows: {ms:.2f} ms")
t0 = time.perf_counter()
for _ in range(5):
    step.enc._stage.upload([s.features for s in step.seqs], dev)
print(f"  upload() host-side return: {(time.perf_counter()-t0)/5*1e3:.2f} ms"); torch.cuda.synchronize()
it = [0]
def e2e():
    step.end_to_end(it[0]); it[0] += 1
ms = tm(e2e, n=10); print(f"end_to_end (two lanes in flight): {ms:.2f} ms/step")
ms = tm(step.resident); print(f"resident eager: {ms:.2f} ms")
ln = step.lanes[0]
def copies():
    hi, d = step.host, ln.d
    d["f1"].copy_(hi["f1"], non_blocking=True); d["f2"].copy_(hi["f2"], non_blocking=True)
    for dc, hc in zip(d["coords"], hi["coords"]): dc.copy_(hc, non_blocking=True)
    for dl, hl in zip(d["eem"], hi["eem"]):
        for k in dl: dl[k].copy_(hl[k], non_blocking=True)
ms = tm(copies); print(f"pinned tensor copies only ({(bench.h2d_bytes(inp)-sum(e.nbytes for e in inp['events']))/1e6:.0f} MB): {ms:.2f} ms")
def d2h():
    ln.out_host.copy_(step.out, non_blocking=True); ln.flow_host.copy_(step.flow, non_blocking=True)
ms = tm(d2h); print(f"D2H results only: {ms:.2f} ms")
