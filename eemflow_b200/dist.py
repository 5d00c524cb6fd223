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
