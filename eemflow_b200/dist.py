"""Batch sharding across GPUs: one process per GPU, independent frame pairs, no data-path collective.

The reference's only multi-GPU mechanism is `torch.nn.DataParallel` splitting dim 0
(train_EEMFlow_HREM.py:116-117).  Event windows and frame pairs are independent, so here each rank
takes a contiguous shard of the batch and runs the whole hot path locally; NCCL (over NVLink /
NVSwitch on the 8xB200 box) is used only to gather the resulting flows in rank order and to reduce
metric accumulators.  The same code runs on the `gloo` backend for the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` for `rank`: pairs i*B/G ... (i+1)*B/G - 1."""
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return lo, hi


def shard(seq, rank: int | None = None, world: int | None = None):
    """Slice a list / tensor along dim 0 for this rank."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(seq), rank, world)
    return seq[lo:hi]


def gather_batch(local: torch.Tensor, total: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather per-rank results `[B_r, ...]` into the rank-ordered global `[B, ...]` on every rank.

    Shards may be ragged (B not divisible by the world size): ranks pad to the largest shard for the
    collective and the padding is cut after it.  Pass `total` (the global batch size) to skip the size
    exchange -- it costs a host synchronisation per call -- and `out` to reuse a result buffer (equal shards).
    The collective is enqueued on the current stream, so it can be overlapped with the next batch by calling
    it under a side stream.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if total is None:
        n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
    else:
        sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    m = max(sizes)
    if all(s == m for s in sizes):
        if out is None:
            out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


def reduce_metrics(acc: torch.Tensor) -> torch.Tensor:
    """Sum metric accumulators (e.g. [epe_sum, n_valid, n_outlier]) over ranks, in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def max_over_ranks(value: float, device: torch.device) -> float:
    """Max of a per-rank scalar (used for device-side timings: the slowest rank defines the step)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _gpu_local_cpus(device_index: int) -> set[int]:
    """CPUs on the GPU's own NUMA node / PCIe root (NVML's ideal affinity), or an empty set when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        return {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
    except Exception:
        return set()


def pin_host_threads(local_rank: int, local_world: int) -> list[int]:
    """Give this rank its own share of the host cores ON ITS GPU'S NUMA NODE (os.sched_setaffinity) and size the host
    thread pools to it.  Call it BEFORE allocating pinned memory: first touch then places the pages next to the GPU's
    PCIe root, so the DMA does not cross the socket interconnect.  Eight ranks that each run an 8-thread staging pool
    on the same 32 cores (torchrun gives no affinity) thrash each other: round 1 measured per-rank host->device
    throughput falling from 44 to 13.5 GB/s.  Returns the cores now owned (empty when the platform offers no affinity call)."""
    if local_world <= 1 or not hasattr(os, "sched_getaffinity"):
        return []
    try:
        avail = sorted(os.sched_getaffinity(0))
        local = sorted(set(avail) & _gpu_local_cpus(local_rank))
        if local:
            # the ranks whose GPUs share this CPU set split it among themselves, in rank order
            sharers = [r for r in range(local_world) if sorted(set(avail) & _gpu_local_cpus(r)) == local]
            pos, cnt = sharers.index(local_rank), len(sharers)
            per = max(1, len(local) // cnt)
            mine = local[pos * per:(pos + 1) * per] or local
        else:
            per = max(1, len(avail) // local_world)
            mine = avail[local_rank * per:(local_rank + 1) * per] or avail
        os.sched_setaffinity(0, mine)
    except (OSError, ValueError):
        return []
    torch.set_num_threads(max(1, len(mine)))
    from . import event_utils
    event_utils.set_staging_threads(max(1, len(mine) - 1))
    return mine


class ResultSink:
    """Delivery of every rank's per-step result (e.g. the final flows `[B_r, 2, H, W]`) to ONE rank.

    The reference gathers results with nn.DataParallel (train_EEMFlow_HREM.py:116-117); evaluation consumes them on
    one process.  Here rank `dst` owns a buffer `[slots, world, *shape]`:

      direct   (peer memory available: torch symmetric memory over NVLink / NVSwitch)  every rank maps dst's buffer
               into its own address space; `slot(k)` is this rank's slice of slot k.  `after_write` / `push` copy the
               result there with a device-to-peer copy on a communication stream (copy engines over NVLink: no
               collective kernel, no SM, nothing competes with the step), overlapped with the next step.  A producer
               may also store into `slot(k)` directly from the kernel that forms the result (`write_through=True`);
               with N - 1 ranks storing into ONE rank that puts dst's inbound link (161 MB per MVSEC step at 8 GPUs
               = 0.21 ms at 770 GB/s) on every producer's critical path, so it is not the default.
      gather   (fallback) `after_write` / `push` enqueue an NCCL gather to dst on a communication stream, overlapped
               with the next step; `before_write` makes the producer wait until the previous gather has read the slot.

    Slots alternate (step k uses slot k % slots) so that dst may consume step i while step i+1 is being produced.
    """

    def __init__(self, shape, dtype, device, dst: int = 0, slots: int = 2, write_through: bool = False):
        assert dist.is_initialized()
        self.write_through = write_through
        self.world, self.rank, self.dst, self.slots = dist.get_world_size(), dist.get_rank(), dst, slots
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.direct = False
        self.peer = None
        self.why = ""
        ok = 0
        if os.environ.get("EEM_RESULT_SINK", "direct") == "direct" and device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                full = (slots, self.world) + self.shape
                self._buf = symm_mem.empty(full, dtype=dtype, device=device)
                self._hdl = symm_mem.rendezvous(self._buf, dist.group.WORLD)
                self.peer = self._hdl.get_buffer(dst, full, dtype)
                ok = 1
            except Exception as e:           # no peer memory on this platform: fall back to NCCL
                self.why = f"{type(e).__name__}: {e}"[:200]
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # all ranks take the same path
        self.direct = bool(flag.item())
        self.cuda = device.type == "cuda"
        self.comm = torch.cuda.Stream(device) if self.cuda else None
        self.read_done = [None] * slots
        self.gathered = None
        if not self.direct:
            self.peer = None
            if self.rank == dst:
                self.gathered = torch.empty((slots, self.world) + self.shape, dtype=dtype, device=device)

    def slot(self, k: int) -> torch.Tensor:
        """This rank's slice of result slot k in dst's memory (direct mode only)."""
        assert self.direct
        return self.peer[k % self.slots, self.rank]

    def buffer(self) -> torch.Tensor | None:
        """On dst: `[slots, world, *shape]`; elsewhere None."""
        if self.rank != self.dst:
            return None
        return self.peer if self.direct else self.gathered

    def before_write(self, k: int) -> None:
        """Call before `local` of slot k is overwritten: waits (on the current stream) until the delivery issued for it
        has read it."""
        if self.cuda and self.read_done[k % self.slots] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.read_done[k % self.slots])

    def after_write(self, k: int, local: torch.Tensor) -> None:
        """`local` has been produced on the current stream: deliver it to dst on the communication stream."""
        if self.direct and self.write_through:
            return                        # the producing kernel stored into slot(k) itself
        if not self.cuda:                 # host tensors (gloo): a synchronous gather
            self._gather(k, local)
            return
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ready)
            if self.direct:
                self.slot(k).copy_(local, non_blocking=True)       # device-to-peer copy over NVLink (copy engine)
            else:
                self._gather(k, local)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.read_done[k % self.slots] = ev

    def push(self, k: int, local: torch.Tensor) -> None:
        """Deliver a result that lives in this rank's own memory."""
        self.before_write(k)
        self.after_write(k, local)

    def _gather(self, k: int, local: torch.Tensor) -> None:
        if self.rank == self.dst:
            dist.gather(local.contiguous(), list(self.gathered[k % self.slots].unbind(0)), dst=self.dst)
        else:
            dist.gather(local.contiguous(), None, dst=self.dst)

    def drain(self) -> None:
        if self.comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)

    def describe(self) -> dict:
        if self.direct:
            mode = ("peer memory over NVLink: stores of the producing kernel (write-through)" if self.write_through else
                    "peer memory over NVLink: device-to-peer copies on a communication stream (copy engines, no collective kernel)")
        else:
            mode = "NCCL gather to one rank on a communication stream"
        return {"mode": mode, "dst": self.dst, "slots": self.slots, "fallback_reason": self.why or None}

    def close(self) -> None:
        self.peer = None
        self._buf = None
        self._hdl = None
