#!/usr/bin/env python
"""Benchmark of the EEMFlow hot path on B200: frame-pairs/s for voxelize + corr + lookup + warp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" is one pass of the hot path over one batch of synthetic frame pairs, BASELINE.json
configs[1] (MVSEC dt1 shape) per GPU:
  * 2*B event windows (260x346, 30 000 events each, [N,4] float64 rows) -> 5-bin voxel grids, normalised
  * CorrBlock on [B,256,36,44] feature maps: 4-level all-pairs pyramid (TF32 tcgen05) + `lookups`
    (default 12 = ERAFT's iterations, model/eraft.py:140) radius-4 window lookups
  * the EEMFlow_cdc op sequence on its 5 pyramid levels at the padded 320x384 size: local 9x9
    correlation (53 kept channels), upsample2d_flow_as, WarpingLayer_no_div, CDC blend, warp, and the 5
    final flow upsamples to 260x346 (model/EEMFlow/EEMFlow+.py:158-234 without the cuDNN convs)
Weak scaling: every rank processes its own B pairs (independent units, no data-path collective);
at N > 1 the per-step result flows are all-gathered over NCCL and a metric accumulator all-reduced.

`value` times device-resident inputs with CUDA events; `e2e` runs the same step through the public
reference-shaped API from HOST buffers (numpy events, pinned feature maps) including H2D and D2H.
`--impl reference` times the CPU oracle port (the reference's own ATen calls, all host threads) on
a bounded sample of the same workload.  Prints ONE JSON line on rank 0.
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

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

H, W, NB = 260, 346, 5                 # MVSEC sensor, a_meshflow / mvsec configs use 5 bins
EVENTS_PER_WINDOW = 30_000             # MVSEC dt1
FH, FW, FD = 36, 44, 256               # ERAFT feature map of the 288x352 padded input
LEVELS, RADIUS = 4, 4
# EEMFlow_cdc pyramid (C, h, w) for the 320x384 padded MVSEC input, coarse -> fine
EEM_LEVELS = [(64, 5, 6), (64, 10, 12), (64, 20, 24), (64, 40, 48), (32, 80, 96)]
# the 53 correlation channels EEMFlow_cdc keeps (model/EEMFlow/EEMFlow+.py:89-97)
EEMFLOW_CDC_INDEX = [0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 21, 22, 23, 24, 26, 28, 29, 30, 31, 32, 33, 34, 36, 38, 39, 40,
                     41, 42, 44, 46, 47, 48, 49, 50, 51, 52, 54, 56, 57, 58, 59, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80]


# Alternative workloads (parity-test shapes of BASELINE.json, not the driver's bench line): selected
# with --workload; the constants above are the default, BASELINE configs[1].
WORKLOADS = {
    "mvsec_dt1": {},
    # configs[2]: HREM dt1 shape -- 720p-class stream, ~10M events/window, 15-bin grids, 92x160 ERAFT
    # feature maps (736x1280 padded / 8), EEMFlow_cdc pyramid of the 768x1280 padded input
    "hrem_dt1": dict(H=720, W=1280, NB=15, EVENTS_PER_WINDOW=10_000_000, FH=92, FW=160,
                     EEM_LEVELS=[(64, 12, 20), (64, 24, 40), (64, 48, 80), (64, 96, 160), (32, 192, 320)], batch=2, cpu_batch=1),
    # configs[3]: HREM dt4 -- 4x longer windows
    "hrem_dt4": dict(H=720, W=1280, NB=15, EVENTS_PER_WINDOW=40_000_000, FH=92, FW=160,
                     EEM_LEVELS=[(64, 12, 20), (64, 24, 40), (64, 48, 80), (64, 96, 160), (32, 192, 320)], batch=1, cpu_batch=1),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mvsec_dt1", choices=sorted(WORKLOADS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="frame pairs per GPU per step (default 32; 2 / 1 for the HREM workloads)")
    ap.add_argument("--lookups", type=int, default=12, help="CorrBlock lookups per pair (ERAFT iterations)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="frame pairs per CPU-baseline sample step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--events-format", default="rows", choices=["rows", "columns"],
                    help="host layout of the events in the end-to-end arm: the reference's [N,4] float64 rows (32 B/event) "
                         "or packed columns as in the HREM .npz files (t f64, x/y int16, p int8 = 13 B/event)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# ------------------------------------------------------------------------------------------------
def make_events(rng, n, h, w):
    t = np.sort(rng.uniform(0.0, 0.05, size=n)) * 1e6
    t -= t[0]
    return np.stack([t, rng.integers(0, w, size=n).astype(np.float64), rng.integers(0, h, size=n).astype(np.float64),
                     2.0 * rng.integers(0, 2, size=n) - 1.0], axis=1)


def make_host_inputs(B: int, lookups: int, seed: int, pin: bool):
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)

    def t(*shape, scale=1.0):
        x = torch.randn(*shape, generator=g) * scale
        return x.pin_memory() if pin else x

    inp = {"events": [make_events(rng, EVENTS_PER_WINDOW, H, W) for _ in range(2 * B)]}
    inp["f1"], inp["f2"] = t(B, FD, FH, FW), t(B, FD, FH, FW)
    base = torch.stack(torch.meshgrid(torch.arange(FH), torch.arange(FW), indexing="ij")[::-1], 0).float()
    inp["coords"] = [(base[None] + t(B, 2, FH, FW, scale=3.0)) for _ in range(lookups)]
    if pin:
        inp["coords"] = [c.pin_memory() for c in inp["coords"]]
    inp["eem"] = []
    for (c, h, w) in EEM_LEVELS:
        inp["eem"].append({"f1": t(B, c, h, w), "f2": t(B, c, h, w), "p1": t(B, 32, h, w), "p2": t(B, 32, h, w),
                           "inter": t(B, 2, h, w, scale=1.5), "mask": torch.sigmoid(t(B, 1, h, w)),
                           "flow": t(B, 2, h, w, scale=2.0)})
    if pin:
        for lv in inp["eem"]:
            for k in lv:
                lv[k] = lv[k].pin_memory()
    return inp


def h2d_bytes(inp, columns: bool = False) -> int:
    ev_bytes = sum(13 * e.shape[0] for e in inp["events"]) if columns else sum(e.nbytes for e in inp["events"])
    n = ev_bytes + inp["f1"].numel() * 4 * 2 + sum(c.numel() * 4 for c in inp["coords"])
    for lv in inp["eem"]:
        n += sum(v.numel() * 4 for v in lv.values())
    return n


# ------------------------------------------------------------------------------------------------
# the step, B200 arm
# ------------------------------------------------------------------------------------------------
class B200Step:
    def __init__(self, inp, dev, lookups):
        import eemflow_b200 as E
        from eemflow_b200 import ops
        from eemflow_b200.correlation import EEMFLOW_CDC_INDEX
        self.E, self.ops, self.index = E, ops, EEMFLOW_CDC_INDEX
        self.dev, self.lookups = dev, lookups
        self.host = inp
        self.B = inp["f1"].shape[0]
        self.target = torch.empty(self.B, 1, H, W, device=dev)
        self.enc = E.EventSequenceToVoxelGrid_Pytorch(NB, gpu=True, gpu_nr=dev.index or 0, normalize=True, forkserver=False)
        self.seqs = [E.EventSequence(None, {"height": H, "width": W}, features=e) for e in inp["events"]]
        self.columns = None       # packed-column copy of the same events (--events-format columns)
        # device-resident copies for the kernel-only number
        ev = np.concatenate(inp["events"], 0)
        self.d_events = torch.from_numpy(ev).to(dev)
        counts = [e.shape[0] for e in inp["events"]]
        self.d_offsets = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int64, device=dev)
        self.max_n = max(counts)
        self.d = {"f1": inp["f1"].to(dev), "f2": inp["f2"].to(dev), "coords": [c.to(dev) for c in inp["coords"]],
                  "eem": [{k: v.to(dev) for k, v in lv.items()} for lv in inp["eem"]]}
        self.out_host = None
        self.flow_host = None
        self.lanes = None

    # ---- the four kernel families of a step; each is also captured as its own CUDA graph ----------
    def fam_voxelize(self):
        self.grids = self.ops.voxelize(self.d_events, self.d_offsets, self.max_n, NB, H, W, normalize=True)

    def fam_corr_pyramid(self, d=None):
        d = d or self.d
        self.blk = self.E.CorrBlock(d["f1"], d["f2"], num_levels=LEVELS, radius=RADIUS, precision="tf32")

    def fam_corr_lookup(self, d=None):
        d = d or self.d
        for c in d["coords"]:
            self.out = self.blk(c)

    def fam_eemflow_ops(self, d=None):
        d = d or self.d
        E = self.E
        flows = []
        lv = d["eem"][0]
        E.correlation_select(lv["f1"], lv["f2"], self.index)
        flow = lv["flow"]
        flows.append(flow)
        for lv in d["eem"][1:]:
            flow_up = E.upsample2d_flow_as(flow.clone(), lv["p1"], mode="bilinear", if_rate=True)
            E.WarpingLayer_no_div()(lv["p2"], flow_up)
            flow_up = E.cdc_blend(flow_up, lv["inter"], lv["mask"])
            f2w = E.warp(lv["f2"], flow_up)
            E.correlation_select(lv["f1"], f2w, self.index)
            flow = flow_up
            flows.append(flow)
        finals = [E.upsample2d_flow_as(f.clone(), self.target, mode="bilinear", if_rate=True) for f in flows]
        self.flow = finals[-1]

    FAMILIES = ("voxelize", "corr_pyramid", "corr_lookup", "eemflow_ops")

    def resident(self):
        """One step with inputs already in HBM; returns (last lookup, final flow)."""
        self.fam_voxelize()
        self.fam_corr_pyramid()
        self.fam_corr_lookup()
        self.fam_eemflow_ops()
        return self.out, self.flow

    class _Lane:
        """Streams, device input buffers, pinned result buffers and a voxel encoder of one in-flight step."""

    def _make_lane(self):
        dev, ln = self.dev, B200Step._Lane()
        ln.d = {"f1": torch.empty_like(self.d["f1"]), "f2": torch.empty_like(self.d["f2"]),
                "coords": [torch.empty_like(c) for c in self.d["coords"]],
                "eem": [{k: torch.empty_like(v) for k, v in lv.items()} for lv in self.d["eem"]]}
        ln.main, ln.h2d, ln.d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ln.enc = self.E.EventSequenceToVoxelGrid_Pytorch(NB, gpu=True, gpu_nr=dev.index or 0, normalize=True, forkserver=False)
        out_shape = (self.B, LEVELS * (2 * RADIUS + 1) ** 2, FH, FW)
        ln.out_host = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
        ln.flow_host = torch.empty((self.B, 2, H, W), dtype=torch.float32, pin_memory=True)
        return ln

    def end_to_end(self, i=0):
        """Same step through the public API from HOST buffers: numpy events, pinned feature maps in,
        last correlation features + final flow out to pinned host memory.

        Two steps are in flight (lane = i % 2), each on its own streams and buffers, so the copies of one step
        overlap the kernels of the other as in any double-buffered serving loop; every step still uploads all of
        its inputs and reads back its results.  Within a lane the pinned tensors go up on a copy stream while the
        host stages the event rows, and the results come back on a third stream."""
        if self.lanes is None:
            self.lanes = [self._make_lane(), self._make_lane()]
        ln = self.lanes[i % 2]
        hi, d = self.host, ln.d
        with torch.cuda.stream(ln.main):
            cur = ln.main
            ln.h2d.wait_stream(cur)               # the lane's previous step is done with these device buffers
            with torch.cuda.stream(ln.h2d):
                d["f1"].copy_(hi["f1"], non_blocking=True)
                d["f2"].copy_(hi["f2"], non_blocking=True)
                for dc, hc in zip(d["coords"], hi["coords"]):
                    dc.copy_(hc, non_blocking=True)
                corr_in = torch.cuda.Event()
                corr_in.record()
            if self.columns is not None:          # stages the events while the copies above are on the link
                ln.enc.voxelize_columns(self.columns, H, W)
            else:
                ln.enc.voxelize_batch(self.seqs)
            with torch.cuda.stream(ln.h2d):       # queued behind the event rows: arrives under voxelize/corr/lookup
                for dl, hl in zip(d["eem"], hi["eem"]):
                    for k in dl:
                        dl[k].copy_(hl[k], non_blocking=True)
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
def reference_step(inp, lookups):
    from oracle import ref_ops as R
    idx = EEMFLOW_CDC_INDEX
    for e in inp["events"]:
        R.voxelize(e, NB, H, W, normalize=True)
    pyr = R.corr_pyramid(inp["f1"], inp["f2"], LEVELS)
    for c in inp["coords"][:lookups]:
        R.corr_lookup(pyr, c, RADIUS)
    B = inp["f1"].shape[0]
    target = torch.empty(B, 1, H, W)
    lv = inp["eem"][0]
    R.correlation(lv["f1"], lv["f2"], 4, index=idx)
    flow = lv["flow"]
    flows = [flow]
    for lv in inp["eem"][1:]:
        flow_up = R.upsample2d_flow_as(flow.clone(), lv["p1"], if_rate=True)
        R.warping_layer_no_div(lv["p2"], flow_up)
        flow_up = R.cdc_blend(flow_up, lv["inter"], lv["mask"])
        f2w = R.warp_exact(lv["f2"], flow_up)
        R.correlation(lv["f1"], f2w, 4, index=idx)
        flow = flow_up
        flows.append(flow)
    return [R.upsample2d_flow_as(f.clone(), target, if_rate=True) for f in flows][-1]


def time_reference(cpu_batch, lookups, steps, warmup, budget_s=25.0):
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_host_inputs(cpu_batch, lookups, seed=1234, pin=False)
    for _ in range(max(1, min(warmup, 2))):
        reference_step(inp, lookups)
    times = []
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        reference_step(inp, lookups)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s and len(times) >= 3:
            break
    return times


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
def lookup_algorithmic_bytes(B):
    """per position 4*L*81 written + 8 coords + 4*sum_l min((2r+2)^2, P_l) volume taps read."""
    P = FH * FW
    taps, h, w = 0, FH, FW
    for _ in range(LEVELS):
        taps += min((2 * RADIUS + 2) ** 2, h * w)
        h, w = h // 2, w // 2
    return B * P * (4 * LEVELS * (2 * RADIUS + 1) ** 2 + 8 + 4 * taps)


def ncu_traffic(kernel, args):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture of
    this very command (profiles/r01/ncu_full_summary_v5.json); only quoted for the workload it was captured on."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01", "ncu_full_summary_v5.json")
    if args.workload != "mvsec_dt1" or args.batch != 32 or not os.path.exists(path):
        return None, None
    try:
        with open(path) as fh:
            d = json.load(fh)[kernel]
        return int(d["dram_rd"] + d["dram_wr"]), "profiles/r01/prof_%s_v5_raw.csv (ncu --set full, one launch)" % kernel
    except (KeyError, ValueError, TypeError):
        return None, None


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    args = parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch is None:
        args.batch = wl.pop("batch", 32)
    if args.cpu_batch is None:
        args.cpu_batch = wl.pop("cpu_batch", 2)
    wl.pop("batch", None), wl.pop("cpu_batch", None)
    globals().update(wl)          # H, W, NB, EVENTS_PER_WINDOW, FH, FW, EEM_LEVELS of the chosen workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"{args.workload}_{H}x{W}_batch{args.batch}_voxelize+corrblock4x4+eemflow_cdc_ops",
              "pairs_per_gpu": args.batch, "windows_per_pair": 2, "events_per_window": EVENTS_PER_WINDOW,
              "voxel": f"{NB}x{H}x{W}", "fmap": f"{FD}x{FH}x{FW}", "corr_levels": LEVELS, "radius": RADIUS,
              "lookups_per_pair": args.lookups, "eemflow_levels": EEM_LEVELS,
              "parallelism": f"batch-sharded x{world}" if world > 1 else "single GPU",
              "l2": "no explicit flush: a step streams > 1 GB (the correlation volume alone is written then gathered 12x) >> 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        times = time_reference(args.cpu_batch, args.lookups, args.steps, args.warmup, budget_s=120.0)
        ms = 1e3 * statistics.mean(times)
        val = args.cpu_batch / statistics.mean(times)
        cores = os.cpu_count() or 1
        sample = f"{args.cpu_batch} frame pairs per step (of the {args.batch}-pair batch), {len(times)} timed steps"
        line = {"impl": "reference", "metric": "frame-pairs/sec (voxelize+corr+lookup+warp)", "value": val, "unit": "frame-pairs/s",
                "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 2), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 event times)", "data": "synthetic",
                "config": dict(config, pairs_per_step=args.cpu_batch),
                "cpu_baseline": {"value": val, "unit": "frame-pairs/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    from eemflow_b200 import _lib
    from eemflow_b200 import dist as edist
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device; there is no CPU fallback"
    rank, world, local_rank = edist.init_from_env("nccl")
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    lib = _lib.lib()

    inp = make_host_inputs(args.batch, args.lookups, seed=100 + rank, pin=True)
    step = B200Step(inp, dev, args.lookups)
    if args.events_format == "columns":
        step.columns = [{"t": np.ascontiguousarray(e[:, 0]), "x": e[:, 1].astype(np.int16), "y": e[:, 2].astype(np.int16),
                         "p": e[:, 3].astype(np.int8)} for e in inp["events"]]

    def eager_step():
        out, flow = step.resident()
        if world > 1:                       # result gather only; nothing on the data path
            edist.gather_batch(flow, total=world * args.batch)
        return out, flow

    for _ in range(max(3, args.warmup)):
        eager_step()
    torch.cuda.synchronize()

    # The step is ~55 short kernels; issued one by one from Python the launch path is as long as the
    # GPU work, so the resident step is captured once into a CUDA graph and replayed (the graph holds
    # exactly the launches of one eager step; buffers live in a shared private pool).  The four kernel
    # families are additionally captured as their own graphs for the per-family split.
    pool = torch.cuda.graph_pool_handle()
    launches_a = lib.eem_launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, pool=pool):
        g_out, g_flow = step.resident()
    launches_per_step = lib.eem_launch_count() - launches_a
    fam_graphs = {}
    for name in B200Step.FAMILIES:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            getattr(step, "fam_" + name)()
        fam_graphs[name] = g

    # N > 1: nothing on the data path is exchanged.  The result gather (flows of all ranks, rank order) and the
    # metric reduction run on a communication stream and overlap the next step's kernels; the timed region ends
    # only after the last gather has completed.  The step is captured twice (same kernels, two sets of output
    # buffers) and the replays alternate, so the collective reads step i's flow while step i+1 writes the other
    # buffer -- no snapshot copy, and nothing but graph launches and event waits on the compute stream.
    graphs, flows = [graph], [g_flow]
    if world > 1:
        from eemflow_b200.eval_utils import flow_error_stats
        comm = torch.cuda.Stream(dev)
        gathered = [torch.empty((world * args.batch,) + tuple(g_flow.shape[1:]), device=dev) for _ in range(2)]
        flow_read = [None, None]
        flow_gt = torch.zeros_like(g_flow).add_(0.5)
        metric_acc = torch.zeros((args.batch, 5), dtype=torch.float64, device=dev)
        with_metrics = os.environ.get("EEM_BENCH_METRICS", "1") != "0"     # timing experiments only

        def step_with_metrics():
            _, flow = step.resident()
            if with_metrics:
                # EPE sums / counts of this rank's pairs, accumulated on the device like the reference's evaluation
                # loop accumulates its AEE sums (test_mvsec.py:291-346); reduced over the ranks once, in drain()
                metric_acc.add_(flow_error_stats(flow_gt, flow))
            return flow

        graphs, flows = [], []
        for _ in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):              # own memory pool: replays overlap with the collective
                f = step_with_metrics()
            graphs.append(g)
            flows.append(f)
        metric_acc.zero_()
    step_no = [0]

    def timed_step():
        k = step_no[0] % len(graphs)
        step_no[0] += 1
        if world == 1:
            graph.replay()
            return
        cur = torch.cuda.current_stream(dev)
        if flow_read[k] is not None:
            cur.wait_event(flow_read[k])            # the gather issued two steps ago has read flows[k]
        graphs[k].replay()
        if os.environ.get("EEM_BENCH_GATHER", "1") != "0":     # timing experiments only
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                edist.gather_batch(flows[k], total=world * args.batch, out=gathered[k])
                flow_read[k] = torch.cuda.Event()
                flow_read[k].record(comm)

    def drain():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(comm)
            edist.reduce_metrics(metric_acc)            # the one metric collective of the run

    for _ in range(3):
        timed_step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        timed_step()
    drain()
    t_stop.record()
    if rank == 0:                       # the steps are only enqueued so far: the device is executing them right now
        for _ in range(4):
            sampler.poke()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    elapsed_ms = edist.max_over_ranks(t_start.elapsed_time(t_stop), dev)
    launches = launches_per_step * args.steps

    # Same K steps again as four family graphs with CUDA events between the replays (events cannot be
    # timed inside one replayed graph): per-family split and the dominant kernel's launch duration.
    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    marks = []
    for _ in range(args.steps):
        marks.append(mark())
        for name in B200Step.FAMILIES:
            fam_graphs[name].replay()
            marks.append(mark())
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    fam = {name: 0.0 for name in B200Step.FAMILIES}
    for s_ in range(args.steps):
        m = marks[5 * s_: 5 * s_ + 5]
        for k_, name in enumerate(B200Step.FAMILIES):
            fam[name] += m[k_].elapsed_time(m[k_ + 1])
    total_fam = sum(fam.values())
    lookup_ms = fam["corr_lookup"] / (args.steps * args.lookups)
    peak, peak_src = peaks()
    algo = lookup_algorithmic_bytes(args.batch)
    achieved = algo / (lookup_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic("corr_lookup_kernel", args)
    roofline = {"kernel": "corr_lookup_kernel<4>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo,
                "avg_launch_ms": lookup_ms, "share_of_step": fam["corr_lookup"] / total_fam,
                "timing": "CUDA events around a graph replay of the 12 lookups, K steps, same inputs as the timed region",
                "family_ms_per_step": {k: v / args.steps for k, v in fam.items()}}

    value = args.batch * world * args.steps / (elapsed_ms * 1e-3)

    # end to end through the public API from host buffers
    e2e = None
    if not args.no_e2e:
        k = max(5, args.steps // 2)
        for i in range(max(6, args.warmup)):
            step.end_to_end(i)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for i in range(k):
            flow = step.end_to_end(i)
            if world > 1:
                torch.cuda.current_stream().wait_stream(step.lanes[i % 2].main)
                edist.gather_batch(flow, total=world * args.batch)
        torch.cuda.synchronize()
        dt = edist.max_over_ranks(time.perf_counter() - t0, dev)
        e2e = {"value": args.batch * world * k / dt, "unit": "frame-pairs/s", "h2d_bytes_per_step": h2d_bytes(inp, args.events_format == "columns"), "events_format": args.events_format,
               "d2h_bytes_per_step": step.d2h_bytes(), "steps": k, "ms_per_step": 1e3 * dt / k,
               "timer": "host perf_counter around synchronize (host staging + H2D + kernels + D2H)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times = time_reference(args.cpu_batch, args.lookups, steps=12, warmup=1, budget_s=20.0)
        cpu = {"value": args.cpu_batch / statistics.mean(times), "unit": "frame-pairs/s", "cores": os.cpu_count() or 1,
               "kind": "port", "sample": f"{args.cpu_batch} frame pairs per step x {len(times)} steps of the same workload, "
                                         f"oracle/ref_ops.py (the reference's ATen calls), torch threads = {torch.get_num_threads()}"}

    if rank == 0:
        line = {"metric": "frame-pairs/sec (voxelize+corr+lookup+warp)", "value": value, "unit": "frame-pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 tensor-core volume, f64 event times)",
                "data": "synthetic", "config": dict(config, launch="CUDA graph replay of one step"), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
