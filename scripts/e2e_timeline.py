"""GPU-side timeline of one end-to-end step (events on each stream, relative to the step start)."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
dev = torch.device("cuda:0")
inp = bench.make_host_inputs(32, 12, 0, pin=True)
step = bench.B200Step(inp, dev, 12)
for _ in range(3):
    step.end_to_end()
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
for it in range(6):
    marks = {}
    def mark(name, stream=None):
        e = E(); e.record(stream or torch.cuda.current_stream()); marks[name] = e
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    cur = torch.cuda.current_stream(dev)
    mark("start")
    hi, d = step.host, step.e2e_d
    step.h2d_stream.wait_stream(cur)
    with torch.cuda.stream(step.h2d_stream):
        d["f1"].copy_(hi["f1"], non_blocking=True); d["f2"].copy_(hi["f2"], non_blocking=True)
        for dc, hc in zip(d["coords"], hi["coords"]): dc.copy_(hc, non_blocking=True)
        mark("corr_in")
    h1 = time.perf_counter()
    step.enc.voxelize_batch(step.seqs)
    h2 = time.perf_counter()
    mark("voxel_done")
    with torch.cuda.stream(step.h2d_stream):
        for dl, hl in zip(d["eem"], hi["eem"]):
            for k in dl: dl[k].copy_(hl[k], non_blocking=True)
        mark("eem_in")
    cur.wait_event(marks["corr_in"])
    step.fam_corr_pyramid(d); step.fam_corr_lookup(d)
    mark("lookup_done")
    out = step.out
    step.d2h_stream.wait_stream(cur)
    with torch.cuda.stream(step.d2h_stream):
        step.out_host.copy_(out, non_blocking=True); mark("out_d2h_done")
    cur.wait_event(marks["eem_in"])
    step.fam_eemflow_ops(d)
    mark("eem_done")
    step.flow_host.copy_(step.flow, non_blocking=True)
    cur.wait_stream(step.d2h_stream)
    mark("end")
    h3 = time.perf_counter()
    torch.cuda.synchronize()
    h4 = time.perf_counter()
    print(f"host: copies enqueued {1e3*(h1-h0):.2f}, staged {1e3*(h2-h0):.2f}, all enqueued {1e3*(h3-h0):.2f}, synced {1e3*(h4-h0):.2f} ms")
    print("gpu :", ", ".join(f"{k} {marks['start'].elapsed_time(v):.2f}" for k, v in marks.items() if k != "start"))
