import sys, time
sys.path.insert(0, ".")
import numpy, torch
import bench
from eemflow_b200 import event_utils as EU
dev = torch.device("cuda:0")
rng = numpy.random.default_rng(0)
inp = bench.make_host_inputs(32, 12, 0, pin=True)
step = bench.B200Step(inp, dev, 12)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 0):
    step.end_to_end()
arrays = [s.features for s in step.seqs]
st = step.enc._stage
for i in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    counts = [int(a.shape[0]) for a in arrays]; total = sum(counts)
    if st.buf is None:
        st.buf = torch.empty((total, 4), dtype=torch.float64, pin_memory=True)
    host = st.buf.numpy()
    t1 = time.perf_counter()
    ev = torch.empty((total, 4), dtype=torch.float64, device=dev)
    t2 = time.perf_counter()
    tasks, pos = [], 0
    for a, n in zip(arrays, counts):
        tasks.append((pos, pos + n, a, 0, n)); pos += n
    def stage(task):
        d0, d1, a, lo, hi = task
        host[d0:d1] = a[lo:hi]
        return d1
    sent = 0; tc = 0.0; tmax = 0.0; tl = time.perf_counter()
    for staged in EU._staging_pool().map(stage, tasks):
        tn = time.perf_counter(); tmax = max(tmax, tn - tl); tl = tn
        if staged - sent >= EU._STAGE_FLUSH_ROWS or staged == total:
            ta = time.perf_counter()
            ev[sent:staged].copy_(st.buf[sent:staged], non_blocking=True)
            tc += time.perf_counter() - ta
            sent = staged
    t3 = time.perf_counter()
    print(f"call {i}: prep {1e3*(t1-t0):.2f} alloc {1e3*(t2-t1):.2f} stage+copy {1e3*(t3-t2):.2f} (copy_ calls {1e3*tc:.2f}, longest wait for a staged chunk {1e3*tmax:.2f}) ms")
