import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from eemflow_b200 import ops
dev = torch.device("cuda:0")
inp = bench.make_host_inputs(32, 12, 0, pin=True)
step = bench.B200Step(inp, dev, 12)
arrays = [s.features for s in step.seqs]
st = step.enc._stage
def T(label, fn, n=6):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        if i: print(f"{label}: host {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t0):.2f} ms")
    return r
ev, off, mx = T("upload", lambda: st.upload(arrays, dev))
T("ops.voxelize", lambda: ops.voxelize(ev, off, mx, bench.NB, bench.H, bench.W, normalize=True))
T("voxelize_batch", lambda: step.enc.voxelize_batch(step.seqs))
print("strict", step.enc.strict)
