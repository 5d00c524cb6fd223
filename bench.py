#!/usr/bin/env python
"""Benchmark of the EEMFlow hot path on B200: frame-pairs/s for voxelize + corr + lookup + warp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workloads a,b,...]

One "step" is one pass of the hot path over one batch of synthetic frame pairs per GPU:
  * 2*B event windows ([N,4] float64 rows) -> voxel grids, normalised
  * CorrBlock on [B,256,h,w] feature maps: 4-level all-pairs pyramid (TF32 tcgen05) + `lookups`
    (default 12 = ERAFT's iterations, model/eraft.py:140) radius-4 window lookups
  * the EEMFlow_cdc op sequence on its 5 pyramid levels at the padded size: local 9x9 correlation
    (53 kept channels), upsample2d_flow_as, WarpingLayer_no_div, CDC blend, warp, and the 5 final flow
    upsamples (model/EEMFlow/EEMFlow+.py:158-234 without the cuDNN convs)
  * the masked end-point-error statistics of the final flow (test_mvsec.py:291-346), accumulated on the device
The headline workload (top-level keys of the JSON line) is BASELINE.json configs[1]: MVSEC dt1 shape, 260x346,
30 000 events per window, 5 bins, batch 32.  The HREM-shaped configs[2] / configs[3] (720x1280, 10 M / 40 M events
per window, 15 bins, 92x160 feature maps) are run in the same invocation and reported under "workloads", each with
its own value / e2e / roofline (dominant kernel family of THAT workload) / cpu_baseline.
Weak scaling: every rank processes its own B pairs (independent units, no data-path collective); at N > 1 the
result flows of every rank are delivered to rank 0 each step (written straight into rank 0's memory over NVLink by
the kernel that produces them when peer memory is available, NCCL gather otherwise) and the metric accumulators
are all-reduced once.  N = 1 and N > 1 time the SAME captured step.

`value` times device-resident inputs with CUDA events; `e2e` runs the same step through the public
reference-shaped API from HOST buffers including H2D and D2H.  `--impl reference` times the CPU oracle port (the
reference's own ATen calls, all host threads) on the same configuration.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FD = 256                               # ERAFT feature channels
LEVELS, RADIUS = 4, 4
# the 53 correlation channels EEMFlow_cdc keeps (model/EEMFlow/EEMFlow+.py:89-97)
EEMFLOW_CDC_INDEX = [0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 21, 22, 23, 24, 26, 28, 29, 30, 31, 32, 33, 34, 36, 38, 39, 40,
                     41, 42, 44, 46, 47, 48, 49, 50, 51, 52, 54, 56, 57, 58, 59, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80]

HREM_LEVELS = [(64, 12, 20), (64, 24, 40), (64, 48, 80), (64, 96, 160), (32, 192, 320)]
WORKLOADS = {
    # configs[1]: MVSEC dt1 shape -- 260x346 sensor, 30 000 events / window, 5 bins, ERAFT feature map of
    # the 288x352 padded input, EEMFlow_cdc pyramid (C, h, w) of the 320x384 padded input
    "mvsec_dt1": dict(H=260, W=346, NB=5, EVENTS_PER_WINDOW=30_000, FH=36, FW=44,
                      EEM_LEVELS=[(64, 5, 6), (64, 10, 12), (64, 20, 24), (64, 40, 48), (32, 80, 96)], batch=32, cpu_batch=32),
    # configs[2]: HREM dt1 shape -- 720p-class stream, ~10M events/window, 15-bin grids, 92x160 ERAFT
    # feature maps (736x1280 padded / 8), EEMFlow_cdc pyramid of the 768x1280 padded input
    "hrem_dt1": dict(H=720, W=1280, NB=15, EVENTS_PER_WINDOW=10_000_000, FH=92, FW=160, EEM_LEVELS=HREM_LEVELS, batch=2, cpu_batch=1),
    # configs[3]: HREM dt4 -- 4x longer windows
    "hrem_dt4": dict(H=720, W=1280, NB=15, EVENTS_PER_WINDOW=40_000_000, FH=92, FW=160, EEM_LEVELS=HREM_LEVELS, batch=1, cpu_batch=1),
}
METRIC = "frame-pairs/sec (voxelize+corr+lookup+warp)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mvsec_dt1", choices=sorted(WORKLOADS), help="headline workload (top-level keys)")
    ap.add_argument("--workloads", default=None,
                    help="comma-separated extra workloads reported under \"workloads\" (default: the HREM shapes when the "
                         "headline is mvsec_dt1; 'none' disables)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="frame pairs per GPU per step of the headline workload")
    ap.add_argument("--lookups", type=int, default=12, help="CorrBlock lookups per pair (ERAFT iterations)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="frame pairs per CPU step (default: the whole batch for mvsec_dt1)")
    ap.add_argument("--corr", default="tf32_f16", choices=["tf32", "tf32_f16", "fp32"], help="CorrBlock precision / storage")
    ap.add_argument("--local-corr", default="tf32", choices=["fp32", "tf32"],
                    help="arithmetic of the local 9x9 correlations: the tcgen05 banded GEMM (TF32 products, fp32 accumulation; the "
                         "same end-to-end gate as the TF32 correlation volume, tests/test_gpu_e2e.py) or the exact FFMA kernel")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[4] instead of the step bench: end-to-end EEMFlow_cdc inference at HREM resolution, global "
                         "batch 1..256 sharded over the ranks, with the CPU path (batch 1) beside it; prints one JSON line")
    ap.add_argument("--sweep-max-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--events-format", default="auto", choices=["auto", "rows", "columns"],
                    help="host layout of the events in the end-to-end arm: the [N,4] float64 rows of EventSequence.features "
                         "(32 B/event; what the MVSEC loader holds: loader/loader_utils.py:44-52) or packed columns as in the HREM "
                         "events{1,2}.npz files (x/y int16, t f64, p int8 = 13 B/event; loader/loader_utils.py:26-37 stacks them "
                         "into rows on the host).  auto: each dataset's own format -- rows for mvsec_*, columns for hrem_*")
    return ap.parse_args()


def workload(name, batch=None, cpu_batch=None):
    wl = SimpleNamespace(name=name, **WORKLOADS[name])
    if batch is not None:
        wl.batch = batch
    if cpu_batch is not None:
        wl.cpu_batch = cpu_batch
    wl.cpu_batch = min(wl.cpu_batch, wl.batch)
    return wl


def config_of(wl, args, world):
    """Identical in both arms (the driver compares them)."""
    return {"workload": f"{wl.name}_{wl.H}x{wl.W}_batch{wl.batch}_voxelize+corrblock4x4+eemflow_cdc_ops",
            "pairs_per_gpu": wl.batch, "windows_per_pair": 2, "events_per_window": wl.EVENTS_PER_WINDOW,
            "voxel": f"{wl.NB}x{wl.H}x{wl.W}", "fmap": f"{FD}x{wl.FH}x{wl.FW}", "corr_levels": LEVELS, "radius": RADIUS,
            "lookups_per_pair": args.lookups, "eemflow_levels": [list(l) for l in wl.EEM_LEVELS],
            "parallelism": f"batch-sharded x{world}" if world > 1 else "single GPU",
            "pairs_per_step_reference_arm": wl.cpu_batch,
            "l2": "no explicit flush: a step streams > 1 GB (the correlation volume alone is written then gathered 12x) >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# ------------------------------------------------------------------------------------------------
def make_events(rng, n, h, w, pin=False):
    """t: sorted uniform stamps in [0, 50 ms) as microseconds relative to the first event (drawn as normalised
    cumulative exponential gaps = the order statistics of n uniforms, O(n) instead of a sort); x, y uniform;
    p in {-1, +1}.  [n, 4] float64 rows, in pinned memory when `pin` (a loader with pin_memory=True)."""
    out = torch.empty((n, 4), dtype=torch.float64, pin_memory=pin).numpy()
    gaps = rng.standard_exponential(n + 1)
    t = np.cumsum(gaps[:n])
    t *= 0.05e6 / (t[-1] + gaps[n])
    t -= t[0]
    out[:, 0] = t
    out[:, 1] = rng.integers(0, w, size=n)
    out[:, 2] = rng.integers(0, h, size=n)
    out[:, 3] = rng.integers(0, 2, size=n)
    out[:, 3] *= 2.0
    out[:, 3] -= 1.0
    return out


class Arena:
    """One pinned host buffer (or one device buffer) carved into tensors, so a step's inputs cross PCIe as a few large
    copies instead of ~70 small ones (each small copy pays a fixed DMA set-up cost: 46 -> 5x GB/s end to end)."""

    def __init__(self, layout, device=None):
        self.layout = layout                      # [(key, shape, offset_bytes)], float32
        total = layout[-1][2] + 4 * int(np.prod(layout[-1][1])) if layout else 0
        self.buf = (torch.empty(total, dtype=torch.uint8, pin_memory=True) if device is None
                    else torch.empty(total, dtype=torch.uint8, device=device))

    @staticmethod
    def plan(named_shapes):
        layout, off = [], 0
        for key, shape in named_shapes:
            layout.append((key, tuple(shape), off))
            off = (off + 4 * int(np.prod(shape)) + 255) // 256 * 256
        return layout

    def views(self):
        return {key: self.buf[off:off + 4 * int(np.prod(shape))].view(torch.float32).view(shape) for key, shape, off in self.layout}


def make_host_inputs(wl, B, lookups, seed, pin):
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)

    def t(*shape, scale=1.0):
        return torch.randn(*shape, generator=g) * scale

    inp = {}
    n = wl.EVENTS_PER_WINDOW
    if pin:     # all windows back to back in ONE pinned buffer (a loader that fills a pinned staging ring)
        ev_all = torch.empty((2 * B * n, 4), dtype=torch.float64, pin_memory=True).numpy()
        inp["events"] = []
        for k in range(2 * B):
            ev_all[k * n:(k + 1) * n] = make_events(rng, n, wl.H, wl.W)
            inp["events"].append(ev_all[k * n:(k + 1) * n])
    else:
        inp["events"] = [make_events(rng, n, wl.H, wl.W) for _ in range(2 * B)]
    corr = {"f1": t(B, FD, wl.FH, wl.FW), "f2": t(B, FD, wl.FH, wl.FW)}
    base = torch.stack(torch.meshgrid(torch.arange(wl.FH), torch.arange(wl.FW), indexing="ij")[::-1], 0).float()
    for k in range(lookups):
        corr[f"coords{k}"] = base[None] + t(B, 2, wl.FH, wl.FW, scale=3.0)
    eem = {}
    for l, (c, h, w) in enumerate(wl.EEM_LEVELS):
        lv = {"f1": t(B, c, h, w), "f2": t(B, c, h, w), "p1": t(B, 32, h, w), "p2": t(B, 32, h, w),
              "inter": t(B, 2, h, w, scale=1.5), "mask": torch.sigmoid(t(B, 1, h, w)), "flow": t(B, 2, h, w, scale=2.0)}
        for k, v in lv.items():
            eem[f"{l}.{k}"] = v
    if pin:     # the float inputs in two pinned arenas: what the correlation needs first, then the EEMFlow-level maps
        for name, group in (("arena_corr", corr), ("arena_eem", eem)):
            arena = Arena(Arena.plan([(k, v.shape) for k, v in group.items()]))
            views = arena.views()
            for k, v in group.items():
                views[k].copy_(v)
                group[k] = views[k]
            inp[name] = arena
    inp["f1"], inp["f2"] = corr["f1"], corr["f2"]
    inp["coords"] = [corr[f"coords{k}"] for k in range(lookups)]
    inp["eem"] = [{k: eem[f"{l}.{k}"] for k in ("f1", "f2", "p1", "p2", "inter", "mask", "flow")} for l in range(len(wl.EEM_LEVELS))]
    inp["flow_gt"] = t(B, 2, wl.H, wl.W, scale=2.0)
    return inp


def h2d_bytes(inp, columns=False):
    ev_bytes = sum(13 * e.shape[0] for e in inp["events"]) if columns else sum(e.nbytes for e in inp["events"])
    n = ev_bytes + inp["f1"].numel() * 4 * 2 + sum(c.numel() * 4 for c in inp["coords"])
    for lv in inp["eem"]:
        n += sum(v.numel() * 4 for v in lv.values())
    return n, ev_bytes


# ------------------------------------------------------------------------------------------------
# the step, B200 arm
# ------------------------------------------------------------------------------------------------
class B200Step:
    FAMILIES = ("voxelize", "corr_pyramid", "corr_lookup", "eemflow_ops", "metrics")

    def __init__(self, wl, inp, dev, lookups, corr):
        import eemflow_b200 as E
        from eemflow_b200 import ops
        from eemflow_b200.eval_utils import flow_error_stats
        self.E, self.ops, self.flow_error_stats = E, ops, flow_error_stats
        self.wl, self.dev, self.lookups, self.corr = wl, dev, lookups, corr
        self.host = inp
        self.B = inp["f1"].shape[0]
        self.target = torch.empty(self.B, 1, wl.H, wl.W, device=dev)
        self.seqs = [E.EventSequence(None, {"height": wl.H, "width": wl.W}, features=e) for e in inp["events"]]
        self.columns = None       # packed-column copy of the same events (--events-format columns)
        # device-resident copies for the kernel-only number
        self.d_events = torch.cat([torch.from_numpy(e).to(dev) for e in inp["events"]], 0)
        counts = [e.shape[0] for e in inp["events"]]
        self.d_offsets = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int64, device=dev)
        self.max_n = max(counts)
        self.d = {"f1": inp["f1"].to(dev), "f2": inp["f2"].to(dev), "coords": [c.to(dev) for c in inp["coords"]],
                  "eem": [{k: v.to(dev) for k, v in lv.items()} for lv in inp["eem"]]}
        self.flow_gt = inp["flow_gt"].to(dev)
        self.metric_acc = torch.zeros((self.B, 5), dtype=torch.float64, device=dev)
        self.flow_out = None      # optional destination of the final flow (e.g. a slot of rank 0's result buffer)
        self.out_host = None
        self.flow_host = None
        self.lanes = None

    # ---- the kernel families of a step; each is also captured as its own CUDA graph ----------
    def fam_voxelize(self):
        wl = self.wl
        self.grids = self.ops.voxelize(self.d_events, self.d_offsets, self.max_n, wl.NB, wl.H, wl.W, normalize=True)

    def fam_voxel_vote(self):
        """K1 alone (zero-init + votes -> raw grid), for the voxelization roofline."""
        wl = self.wl
        self.raw = self.ops.voxelize(self.d_events, self.d_offsets, self.max_n, wl.NB, wl.H, wl.W, normalize=False)

    def fam_corr_pyramid(self, d=None):
        d = d or self.d
        self.blk = self.E.CorrBlock(d["f1"], d["f2"], num_levels=LEVELS, radius=RADIUS, precision=self.corr)

    def fam_corr_lookup(self, d=None):
        d = d or self.d
        for c in d["coords"]:
            self.out = self.blk(c)

    def fam_eemflow_ops(self, d=None):
        """The EEMFlow_cdc call sequence (EEMFlow+.py:158-234) on synthetic feature maps.  Like the reference, the flow of a
        level is scaled IN PLACE by the upsample2d_flow_as of the next level before its own final upsampling; only the
        coarsest flow (a persistent bench input) is cloned so that steps do not compound."""
        d = d or self.d
        E = self.E
        index = EEMFLOW_CDC_INDEX
        lv = d["eem"][0]
        E.correlation_select(lv["f1"], lv["f2"], index)
        flow = lv["flow"].clone()
        flows, pre = [flow], []
        for lv in d["eem"][1:]:
            # as eemflow_b200.models.EEMFlow_cdc issues them: upsample2d_flow_as + WarpingLayer_no_div in one launch,
            # CDC blend + warp in one launch, the in-place scaling of the coarse flow deferred to the final upsampling
            flow_up, _, sc = E.upsample_warp_no_div(flow, lv["p1"], lv["p2"])
            flow_up, f2w = E.blend_warp(flow_up, lv["inter"], lv["mask"], lv["f2"])
            E.correlation_select(lv["f1"], f2w, index)
            pre.append(sc)
            flow = flow_up
            flows.append(flow)
        finals = E.upsample2d_flows_as(flows, self.target, mode="bilinear", if_rate=True, out_last=self.flow_out, pre_scales=pre + [None])
        self.flow = finals[-1]

    def fam_metrics(self):
        # EPE sums / counts of this rank's pairs, accumulated on the device like the reference's evaluation loop
        # accumulates its AEE sums (test_mvsec.py:291-346); reduced over the ranks once at the end of the run
        self.metric_acc.add_(self.flow_error_stats(self.flow_gt, self.flow))

    def resident(self):
        """One step with inputs already in HBM; returns (last lookup, final flow).

        The three chains of a step are independent of each other -- event windows -> voxel grids; feature maps ->
        pyramid -> 12 lookups; EEMFlow-level maps -> flows -> metrics -- so (unless EEM_BENCH_SERIAL=1) they are issued
        on three streams and become three concurrent branches of the captured step graph: the small latency-bound
        kernels of one chain fill the issue slots and SMs the other chains leave idle.  Same kernels, same results."""
        if os.environ.get("EEM_BENCH_SERIAL", "0") == "1":
            self.fam_voxelize()
            self.fam_corr_pyramid()
            self.fam_corr_lookup()
            self.fam_eemflow_ops()
            self.fam_metrics()
            return self.out, self.flow
        if getattr(self, "side", None) is None:
            # EEM_BENCH_PRIO=1 gives the longest chain (pyramid + lookups) a high-priority stream; measured slightly
            # SLOWER (0.962 vs 0.945 ms/step), so all chains run at the same priority by default
            prio = -1 if os.environ.get("EEM_BENCH_PRIO", "0") == "1" else 0
            self.side = (torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev, priority=prio))
        cur = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        s1, s2, s0 = self.side
        for st in (s0, s1, s2):
            st.wait_event(fork)
        with torch.cuda.stream(s0):
            self.fam_corr_pyramid()
            self.fam_corr_lookup()
        with torch.cuda.stream(s1):
            self.fam_voxelize()
        with torch.cuda.stream(s2):
            self.fam_eemflow_ops()
            self.fam_metrics()
        for st in (s0, s1, s2):
            cur.wait_stream(st)
        return self.out, self.flow

    class _Lane:
        """Streams, device input buffers, pinned result buffers and a voxel encoder of one in-flight step."""

    def _make_lane(self):
        dev, wl, ln = self.dev, self.wl, B200Step._Lane()
        hi = self.host
        # device mirrors of the two pinned input arenas: one large copy each per step
        ln.dev_corr, ln.dev_eem = Arena(hi["arena_corr"].layout, dev), Arena(hi["arena_eem"].layout, dev)
        vc, ve = ln.dev_corr.views(), ln.dev_eem.views()
        ln.d = {"f1": vc["f1"], "f2": vc["f2"], "coords": [vc[f"coords{k}"] for k in range(self.lookups)],
                "eem": [{k: ve[f"{l}.{k}"] for k in ("f1", "f2", "p1", "p2", "inter", "mask", "flow")} for l in range(len(wl.EEM_LEVELS))]}
        ln.main, ln.h2d, ln.d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ln.enc = self.E.EventSequenceToVoxelGrid_Pytorch(wl.NB, gpu=True, gpu_nr=dev.index or 0, normalize=True, forkserver=False)
        out_shape = (self.B, LEVELS * (2 * RADIUS + 1) ** 2, wl.FH, wl.FW)
        ln.out_host = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
        ln.flow_host = torch.empty((self.B, 2, wl.H, wl.W), dtype=torch.float32, pin_memory=True)
        return ln

    N_LANES = 3

    def end_to_end(self, i=0):
        """Same step through the public API from HOST buffers: numpy events, pinned feature maps in,
        last correlation features + final flow out to pinned host memory.

        N_LANES steps are in flight (lane = i % N_LANES), each on its own streams and buffers, so the copies of one
        step overlap the kernels of the others as in any multi-buffered serving loop; every step still uploads all of
        its inputs and reads back its results.  Within a lane the pinned tensors go up on a copy stream while the
        host stages the event rows, and the results come back on a third stream."""
        if self.lanes is None:
            self.lanes = [self._make_lane() for _ in range(self.N_LANES)]
        ln = self.lanes[i % self.N_LANES]
        hi, d, wl = self.host, ln.d, self.wl
        with torch.cuda.stream(ln.main):
            cur = ln.main
            ln.h2d.wait_stream(cur)               # the lane's previous step is done with these device buffers
            with torch.cuda.stream(ln.h2d):
                ln.dev_corr.buf.copy_(hi["arena_corr"].buf, non_blocking=True)      # f1, f2, all lookup coordinates
                corr_in = torch.cuda.Event()
                corr_in.record()
            if self.columns is not None:          # stages the events while the copies above are on the link
                ln.enc.voxelize_columns(self.columns, wl.H, wl.W)
            else:
                ln.enc.voxelize_batch(self.seqs)
            with torch.cuda.stream(ln.h2d):       # queued behind the event rows: arrives under voxelize/corr/lookup
                ln.dev_eem.buf.copy_(hi["arena_eem"].buf, non_blocking=True)        # the EEMFlow-level maps
                eem_in = torch.cuda.Event()
                eem_in.record()
            cur.wait_event(corr_in)
            self.fam_corr_pyramid(d)
            self.fam_corr_lookup(d)
            out = self.out
            ln.d2h.wait_stream(cur)
            with torch.cuda.stream(ln.d2h):
                ln.out_host.copy_(out, non_blocking=True)
                out.record_stream(ln.d2h)
            cur.wait_event(eem_in)
            self.fam_eemflow_ops(d)
            self.fam_metrics()
            ln.flow_host.copy_(self.flow, non_blocking=True)
            cur.wait_stream(ln.d2h)
            self.flow.record_stream(cur)
        self.out_host, self.flow_host = ln.out_host, ln.flow_host
        return self.flow

    def e2e_sync(self):
        for ln in self.lanes or ():
            ln.main.synchronize()

    def d2h_bytes(self):
        return (self.out_host.numel() + self.flow_host.numel()) * 4 if self.out_host is not None else 0


# ------------------------------------------------------------------------------------------------
# the step, CPU reference arm (oracle port: the reference's own ATen calls)
# ------------------------------------------------------------------------------------------------
def reference_step(wl, inp, lookups):
    from oracle import ref_ops as R
    idx = EEMFLOW_CDC_INDEX
    for e in inp["events"]:
        R.voxelize(e, wl.NB, wl.H, wl.W, normalize=True)
    pyr = R.corr_pyramid(inp["f1"], inp["f2"], LEVELS)
    for c in inp["coords"][:lookups]:
        R.corr_lookup(pyr, c, RADIUS)
    B = inp["f1"].shape[0]
    target = torch.empty(B, 1, wl.H, wl.W)
    lv = inp["eem"][0]
    R.correlation(lv["f1"], lv["f2"], 4, index=idx)
    flow = lv["flow"].clone()
    flows = [flow]
    for lv in inp["eem"][1:]:
        flow_up = R.upsample2d_flow_as(flow, lv["p1"], if_rate=True)
        R.warping_layer_no_div(lv["p2"], flow_up)
        flow_up = R.cdc_blend(flow_up, lv["inter"], lv["mask"])
        f2w = R.warp_exact(lv["f2"], flow_up)
        R.correlation(lv["f1"], f2w, 4, index=idx)
        flow = flow_up
        flows.append(flow)
    final = [R.upsample2d_flow_as(f, target, if_rate=True) for f in flows][-1]
    for b in range(B):      # Test.flow_error per sample, dense evaluation (test_mvsec.py:291-346)
        R.flow_error(inp["flow_gt"][b:b + 1], final[b:b + 1], None, False, "dense")
    return final


def time_reference(wl, lookups, steps, warmup, budget_s):
    """Times `steps` CPU steps after `warmup` untimed ones; stops early only when the time budget is spent."""
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_host_inputs(wl, wl.cpu_batch, lookups, seed=1234, pin=False)
    t_all = time.perf_counter()
    done_warm = 0
    for _ in range(warmup):
        reference_step(wl, inp, lookups)
        done_warm += 1
        if time.perf_counter() - t_all > 0.4 * budget_s:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_step(wl, inp, lookups)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    return times, done_warm


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every ~2 ms (a 25 ms timed
    region gets ~10 samples); `nvidia-smi -lms` as the fallback when the NVML bindings are missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, threading.Event(), None
        self.sm, self.mx, self.reasons = [], [], set()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.handle))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def poke(self):
        """One sample taken synchronously by the caller (the main thread calls this right after enqueueing the timed
        steps and the stop event, while the GPU is still executing them, so there are under-load samples even if the polling thread
        did not get the GIL during the few milliseconds of the timed region)."""
        n = self.nvml
        if n is None:
            return
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.sm:
                return None
            return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for k, nme in enumerate(self.NAMES):
                if r[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# roofline of the dominant kernel family (DESIGN.md section 4 defines the algorithmic bytes)
# ------------------------------------------------------------------------------------------------
def lookup_algorithmic_bytes(wl, B, corr):
    """K5, per position: 4*L*81 written + 8 coords + s*sum_l min((2r+2)^2, P_l) volume taps read, s = bytes per
    stored volume element (4 for the f32 pyramid, 2 for the fp16 working pyramid)."""
    s = 2 if corr == "tf32_f16" else 4
    P = wl.FH * wl.FW
    taps, h, w = 0, wl.FH, wl.FW
    for _ in range(LEVELS):
        taps += min((2 * RADIUS + 2) ** 2, h * w)
        h, w = h // 2, w // 2
    return B * P * (4 * LEVELS * (2 * RADIUS + 1) ** 2 + 8 + s * taps)


def voxel_algorithmic_bytes(wl, B):
    """K1, per window: 32*N event-row bytes read + 4*nb*H*W grid bytes written once."""
    return 2 * B * (32 * wl.EVENTS_PER_WINDOW + 4 * wl.NB * wl.H * wl.W)


def ncu_traffic(kernel, wl, B, corr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture of this
    very command (profiles/r0N/ncu_full_summary*.json); only quoted for the workload/batch it was captured on."""
    for rel in ("profiles/r02/ncu_full_summary.json", "profiles/r01/ncu_full_summary_v5.json"):
        path = ROOT / rel
        if not path.exists():
            continue
        try:
            d = json.loads(path.read_text())
            meta = d.get("_meta", {"workload": "mvsec_dt1", "batch": 32, "corr": "tf32"})
            if meta.get("workload") != wl.name or meta.get("batch") != B or meta.get("corr", "tf32") != corr:
                continue
            k = d[kernel]
            return int(k["dram_rd"] + k["dram_wr"]), f"{rel} [{kernel}] (ncu --set full, one launch)"
        except (KeyError, ValueError, TypeError):
            continue
    return None, None


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# one workload on the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(wl, args, rank, world, dev, steps, warmup, do_e2e, do_cpu, sample_clocks):
    from eemflow_b200 import _lib
    from eemflow_b200 import dist as edist
    lib = _lib.lib()
    B = wl.batch
    inp = make_host_inputs(wl, B, args.lookups, seed=100 + rank, pin=True)
    step = B200Step(wl, inp, dev, args.lookups, args.corr)
    events_format = args.events_format if args.events_format != "auto" else ("columns" if wl.name.startswith("hrem") else "rows")
    if events_format == "columns":
        def pinned(a, dtype):          # like the rows: page-locked host memory (a loader with pin_memory=True)
            t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
            t.numpy()[...] = a
            return t.numpy()
        step.columns = [{"t": pinned(e[:, 0], torch.float64), "x": pinned(e[:, 1], torch.int16), "y": pinned(e[:, 2], torch.int16),
                         "p": pinned(e[:, 3], torch.int8)} for e in inp["events"]]

    # N > 1: nothing on the data path is exchanged.  The result flows go to rank 0 every step, on a communication
    # stream that overlaps the next step: a device-to-peer copy into rank 0's result buffer (peer memory over NVLink,
    # copy engines -- no collective kernel), or an NCCL gather when peer memory cannot be set up.  The step is captured
    # twice (two sets of output buffers) and the replays alternate, so step i's flows are read by the delivery while
    # step i+1 writes the other set.  Metric accumulators are reduced once, at the end of the run.
    write_through = os.environ.get("EEM_RESULT_WRITE_THROUGH", "0") == "1"       # timing experiments: stores from the producing kernel
    sink = edist.ResultSink((B, 2, wl.H, wl.W), torch.float32, dev, dst=0, slots=2, write_through=write_through) if world > 1 else None

    for _ in range(max(3, warmup)):
        step.resident()
    torch.cuda.synchronize()

    # The step is ~55 short kernels; issued one by one from Python the launch path is as long as the GPU work, so
    # the resident step is captured once into a CUDA graph and replayed (the graph holds exactly the launches of one
    # eager step).  The kernel families are additionally captured as their own graphs for the per-family split.
    n_graphs = 2 if world > 1 else 1
    graphs, flows = [], []
    launches_per_step = 0
    for k in range(n_graphs):
        step.flow_out = sink.slot(k) if (sink is not None and sink.direct and sink.write_through) else None
        launches_a = lib.eem_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            _, f = step.resident()
        launches_per_step = lib.eem_launch_count() - launches_a
        graphs.append(g)
        flows.append(f)
    step.flow_out = None
    pool = torch.cuda.graph_pool_handle()
    fam_graphs = {}
    for name in B200Step.FAMILIES + ("voxel_vote",):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            getattr(step, "fam_" + name)()
        fam_graphs[name] = g
    step.metric_acc.zero_()
    step_no = [0]

    def timed_step():
        k = step_no[0] % n_graphs
        step_no[0] += 1
        if sink is not None:
            sink.before_write(k)
        graphs[k].replay()
        if sink is not None:
            sink.after_write(k, flows[k])

    def drain():
        if sink is not None:
            sink.drain()
            edist.reduce_metrics(step.metric_acc)            # the one metric collective of the run

    for _ in range(3):
        timed_step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(dev.index or 0) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
    torch.cuda.synchronize()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(steps):
        timed_step()
    drain()
    t_stop.record()
    if sampler:                         # the steps are only enqueued so far: the device is executing them right now
        for _ in range(4):
            sampler.poke()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    elapsed_ms = edist.max_over_ranks(t_start.elapsed_time(t_stop), dev)
    launches = launches_per_step * steps
    gather_verified = None
    if sink is not None:
        # outside the timed region: rank 0's result buffer must hold every rank's last flows (checksums travel by NCCL)
        mine = torch.stack([f.double().sum() for f in flows])
        sums = [torch.zeros_like(mine) for _ in range(world)]
        torch.distributed.all_gather(sums, mine)
        buf = sink.buffer()
        if rank == 0 and buf is not None:
            gather_verified = all(abs(buf[k, r].double().sum().item() - sums[r][k].item()) <= 1e-6 * max(1.0, abs(sums[r][k].item()))
                                  for k in range(n_graphs) for r in range(world))

    # Same K steps again as family graphs with CUDA events between the replays (events cannot be timed inside one
    # replayed graph): per-family split and the dominant kernel's launch duration.
    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    names = B200Step.FAMILIES + ("voxel_vote",)
    marks = []
    for _ in range(steps):
        marks.append(mark())
        for name in names:
            fam_graphs[name].replay()
            marks.append(mark())
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    nm = len(names) + 1
    fam = {name: 0.0 for name in names}
    for s_ in range(steps):
        m = marks[nm * s_: nm * s_ + nm]
        for k_, name in enumerate(names):
            fam[name] += m[k_].elapsed_time(m[k_ + 1])
    fam_ms = {k: v / steps for k, v in fam.items()}
    vote_ms = fam_ms.pop("voxel_vote")
    total_fam = sum(fam_ms.values())
    peak, peak_src = peaks()
    dominant = max(fam_ms, key=fam_ms.get)
    # the roofline is quoted for the dominant family's kernel when that family is one kernel (voxelize -> K1 vote
    # pass, corr_lookup -> K5); for the multi-kernel families the larger of the two single-kernel families is used
    if dominant not in ("voxelize", "corr_lookup"):
        dominant = "voxelize" if fam_ms["voxelize"] >= fam_ms["corr_lookup"] else "corr_lookup"
    if dominant == "corr_lookup":
        kernel = "corr_lookup_packed_ws_kernel<4>" if args.corr == "tf32_f16" else "corr_lookup_kernel<4>"
        ncu_key = "corr_lookup_packed_ws_kernel" if args.corr == "tf32_f16" else "corr_lookup_kernel"
        launch_ms = fam_ms["corr_lookup"] / args.lookups
        algo = lookup_algorithmic_bytes(wl, B, args.corr)
        timing = f"CUDA events around a graph replay of the {args.lookups} lookups, K steps, same inputs as the timed region"
    else:
        kernel, ncu_key = "K1 voxel vote (eem_voxelize, normalize=0: zero-init + votes -> raw grid)", "voxel_vote"
        launch_ms = vote_ms
        algo = voxel_algorithmic_bytes(wl, B)
        timing = "CUDA events around a graph replay of one un-normalised eem_voxelize call over the step's 2B windows, K steps"
    achieved = algo / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(ncu_key, wl, B, args.corr)
    roofline = {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo, "avg_launch_ms": launch_ms,
                "share_of_step": fam_ms[dominant] / total_fam, "timing": timing,
                "family_ms_per_step": fam_ms, "voxel_vote_ms": vote_ms,
                "voxel_vote_frac_of_hbm_peak": voxel_algorithmic_bytes(wl, B) / (vote_ms * 1e-3) / 1e9 / peak}

    value = B * world * steps / (elapsed_ms * 1e-3)

    # end to end through the public API from host buffers
    e2e = None
    if do_e2e:
        k = max(5, steps // 2)
        for i in range(max(2 * B200Step.N_LANES, warmup)):
            step.end_to_end(i)
        # three repetitions of k steps; the MEDIAN repetition is reported (one repetition is ~70 ms of host-timed work,
        # short enough for a single scheduling hiccup of the host to double it)
        reps = []
        for _ in range(3):
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            t0 = time.perf_counter()
            for i in range(k):
                flow = step.end_to_end(i)
                if sink is not None:
                    torch.cuda.current_stream().wait_stream(step.lanes[i % B200Step.N_LANES].main)
                    sink.push(i % 2, flow)
            if sink is not None:
                sink.drain()
            torch.cuda.synchronize()
            reps.append(edist.max_over_ranks(time.perf_counter() - t0, dev))
        dt = statistics.median(reps)
        total_b, ev_b = h2d_bytes(inp, events_format == "columns")
        e2e = {"value": B * world * k / dt, "unit": "frame-pairs/s", "h2d_bytes_per_step": total_b,
               "h2d_event_bytes_per_step": ev_b, "events_format": events_format,
               "d2h_bytes_per_step": step.d2h_bytes(), "steps": k, "ms_per_step": 1e3 * dt / k,
               "h2d_gbs_per_rank": total_b / (dt / k) / 1e9, "repetitions_ms_per_step": [1e3 * r / k for r in reps],
               "timer": "host perf_counter around synchronize (host staging + H2D + kernels + D2H)"}

    cpu = None
    if do_cpu and rank == 0 and world == 1:
        times, _ = time_reference(wl, args.lookups, steps=12, warmup=1, budget_s=25.0)
        cpu = {"value": wl.cpu_batch / statistics.mean(times), "unit": "frame-pairs/s", "cores": os.cpu_count() or 1,
               "kind": "port", "sample": f"{wl.cpu_batch} frame pairs per step x {len(times)} steps of the same workload, "
                                         f"oracle/ref_ops.py (the reference's ATen calls), torch threads = {torch.get_num_threads()}"}

    res = {"value": value, "ms_per_step": elapsed_ms / steps, "steps": steps, "warmup": max(3, warmup), "clocks": clocks,
           "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "result_gather": None if sink is None else dict(sink.describe(), verified=gather_verified)}
    if sink is not None:
        sink.close()
    del step, graphs, fam_graphs, inp
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: end-to-end EEMFlow inference sweep, batch 1..256 at HREM resolution, 1..8 GPUs
# ------------------------------------------------------------------------------------------------
def cpu_model_forward(model, events1, events2):
    """The EEMFlow_cdc forward (model/EEMFlow/EEMFlow+.py:158-234) on the CPU with the model's own (ATen CPU) convolutions
    and the oracle's restatements of the hot-path ops -- the reference's CPU path for the same weights."""
    from oracle import ref_ops as R
    idx = EEMFLOW_CDC_INDEX
    image1, _ = R.input_pad(events1, events1.shape, mode="chairs", eval_pad_rate=64)
    image2, _ = R.input_pad(events2, events2.shape, mode="chairs", eval_pad_rate=64)
    p1, p2 = model._pyramid(image1), model._pyramid(image2)
    flows = {}
    f16, f26 = p1[6], p2[6]
    flow7_up = torch.zeros(f16.size(0), 2, f16.size(2), f16.size(3))
    cv = R.correlation(f16, f26, 4, index=idx)
    flows[6] = model.decoder6(torch.cat([cv, model.rconv6(f16), flow7_up], 1))
    cdc = model.cdc_model
    for lvl in (5, 4, 3, 2):
        a, b = p1[lvl], p2[lvl]
        proj = model.conv_1x1[lvl]
        fa, fb = proj(a), proj(b)
        flow_init = R.upsample2d_flow_as(flows[lvl + 1], fa, if_rate=True)
        x_out = cdc.dense_estimator_mask(torch.cat((fa, R.warping_layer_no_div(fb, flow_init)), dim=1))
        flow_up = R.cdc_blend(flow_init, x_out[:, :2].contiguous(), torch.sigmoid(x_out[:, 2:3]).contiguous())
        cv = R.correlation(a, R.warp_exact(b, flow_up), 4, index=idx)
        feat = getattr(model, f"rconv{lvl}")(a)
        flows[lvl] = getattr(model, f"decoder{lvl}")(torch.cat([cv, feat, flow_up], 1)) + flow_up
    return [R.upsample2d_flow_as(flows[lvl], events1, if_rate=True) for lvl in (6, 5, 4, 3, 2)]


def run_sweep(args):
    from eemflow_b200 import dist as edist
    from eemflow_b200.models import EEMFlow_cdc
    import eemflow_b200
    assert torch.cuda.is_available(), "bench.py --sweep needs a CUDA device"
    eemflow_b200.set_local_corr_precision(args.local_corr)
    rank, world, local_rank = edist.init_from_env("nccl")
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    nb, h, w = 15, 720, 1280
    micro = 32                                  # a rank runs its shard in micro-batches of at most 32 pairs
    torch.manual_seed(0)
    model = EEMFlow_cdc(None, groups=3, n_first_channels=nb).to(dev).eval()
    model.change_imagesize((h, w))
    torch.backends.cudnn.benchmark = True
    rows = []
    B = 1
    while B <= args.sweep_max_batch:
        lo, hi = edist.shard_bounds(B, rank, world)
        mine = hi - lo
        chunks = [min(micro, mine - k) for k in range(0, mine, micro)]
        inputs = [(torch.randn(c, nb, h, w, device=dev), torch.randn(c, nb, h, w, device=dev)) for c in sorted(set(chunks))]
        by_size = {v1.shape[0]: (v1, v2) for v1, v2 in inputs}

        def forward_all():
            with torch.no_grad():
                for c in chunks:
                    v1, v2 = by_size[c]
                    model(events1=v1, events2=v2)

        for _ in range(2):
            forward_all()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            forward_all()
            b.record()
            torch.cuda.synchronize()
            ts.append(edist.max_over_ranks(a.elapsed_time(b), dev))
        ms = statistics.median(ts)
        rows.append({"global_batch": B, "pairs_per_rank_max": -(-B // world), "ms_per_forward": ms, "pairs_per_s": B / ms * 1e3,
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
        del inputs, by_size
        torch.cuda.empty_cache()
        B *= 2
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cm = EEMFlow_cdc(None, groups=3, n_first_channels=nb).eval()
        cm.load_state_dict(model.state_dict())
        v1, v2 = torch.randn(1, nb, h, w), torch.randn(1, nb, h, w)
        with torch.no_grad():
            cpu_model_forward(cm, v1, v2)
            t0 = time.perf_counter()
            n = 0
            while n < 3 and (n == 0 or time.perf_counter() - t0 < 20.0):
                cpu_model_forward(cm, v1, v2)
                n += 1
            dt = (time.perf_counter() - t0) / n
        cpu = {"value": 1.0 / dt, "unit": "frame-pairs/s", "cores": os.cpu_count() or 1, "kind": "port", "ms_per_forward": 1e3 * dt,
               "sample": f"batch 1, {n} forwards, the model's ATen CPU convolutions + oracle/ref_ops.py hot-path ops"}
    if rank == 0:
        print(json.dumps({"metric": "frame-pairs/sec (end-to-end EEMFlow_cdc inference, HREM 720x1280, 15 bins)", "unit": "frame-pairs/s",
                          "n_gpus": world, "higher_is_better": True, "scaling": "strong",
                          "dtype": "f32" + (" (tf32 tensor-core local correlation)" if args.local_corr == "tf32" else ""), "data": "synthetic",
                          "config": {"workload": "eemflow_cdc_inference_sweep_hrem_720x1280", "model": "EEMFlow_cdc(groups=3), random init",
                                     "input": "voxel grids resident on the device", "micro_batch": micro,
                                     "parallelism": f"batch-sharded x{world}"},
                          "sweep": rows, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse_args()
    if args.sweep:
        return run_sweep(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    head = workload(args.workload, args.batch, args.cpu_batch)
    if args.workloads is None:
        extra = ["hrem_dt1", "hrem_dt4"] if args.workload == "mvsec_dt1" else []
    else:
        extra = [w for w in args.workloads.split(",") if w and w != "none"]
    extra = [w for w in extra if w != args.workload]
    dtype = {"tf32": "f32 (tf32 tensor-core volume, f64 event times)", "fp32": "f32 (f64 event times)",
             "tf32_f16": "f32 (tf32 tensor-core volume stored as fp16, f64 event times)"}[args.corr]
    if args.local_corr == "tf32":
        dtype = dtype.replace("f64 event times", "tf32 tensor-core local correlation, f64 event times")

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1

        def ref_line(wl, steps, warmup, budget):
            times, done_warm = time_reference(wl, args.lookups, steps, warmup, budget_s=budget)
            val = wl.cpu_batch / statistics.mean(times)
            sample = f"{wl.cpu_batch} frame pairs per step (of the {wl.batch}-pair batch), {len(times)} timed steps"
            return {"value": val, "unit": "frame-pairs/s", "steps": len(times), "warmup": done_warm,
                    "ms_per_step": 1e3 * statistics.mean(times), "config": config_of(wl, args, world),
                    "cpu_baseline": {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port", "sample": sample},
                    "e2e": {"value": val, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

        r = ref_line(head, args.steps, args.warmup, 240.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frame-pairs/s", "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 event times)", "data": "synthetic",
                "config": r["config"], "cpu_baseline": r["cpu_baseline"], "e2e": r["e2e"]}
        if extra:
            line["workloads"] = {name: ref_line(workload(name), min(args.steps, 3), 1, 60.0) for name in extra}
        print(json.dumps(line), flush=True)
        return

    from eemflow_b200 import dist as edist
    import eemflow_b200
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device; there is no CPU fallback"
    eemflow_b200.set_local_corr_precision(args.local_corr)
    rank, world, local_rank = edist.init_from_env("nccl")
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    cores = edist.pin_host_threads(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    if world > 1:
        print(f"[bench] rank {rank}: {len(cores)} host cores {cores[:4]}..{cores[-1:] if cores else []}", file=sys.stderr, flush=True)

    r = run_b200(head, args, rank, world, dev, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline, True)
    subs = {}
    for name in extra:
        wl = workload(name)
        s = run_b200(wl, args, rank, world, dev, max(3, min(args.steps, 10)), min(args.warmup, 3), not args.no_e2e,
                     not args.no_cpu_baseline, False)
        s["config"] = config_of(wl, args, world)
        s["unit"] = "frame-pairs/s"
        subs[name] = s

    if rank == 0:
        line = {"metric": METRIC, "value": r["value"], "unit": "frame-pairs/s", "n_gpus": world,
                "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": dict(config_of(head, args, world)),
                "launch": "CUDA graph replay of one step; its three independent chains (voxelize | pyramid + lookups | EEMFlow ops + metrics) are concurrent branches of the graph"
                          if os.environ.get("EEM_BENCH_SERIAL", "0") != "1" else "CUDA graph replay of one step, kernels in series",
                "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
                "roofline": r["roofline"], "cpu_baseline": r["cpu_baseline"], "result_gather": r["result_gather"]}
        if subs:
            line["workloads"] = subs
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
