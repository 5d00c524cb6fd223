"""CPU tests of the multi-GPU host logic with world_size 2 on the gloo backend (no GPU needed):
contiguous batch shards, rank-ordered result gather (even and ragged), metric reduction, max-over-ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eemflow_b200 import dist as edist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = edist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    full = torch.arange(total * 6, dtype=torch.float32).view(total, 2, 3)
    local = edist.shard(full)
    lo, hi = edist.shard_bounds(total, rank, world)
    assert torch.equal(local, full[lo:hi])
    got = edist.gather_batch(local * 1.0, total=total)
    got2 = edist.gather_batch(local * 1.0)                     # sizes exchanged by all_gather
    acc = edist.reduce_metrics(torch.tensor([float(rank + 1), 10.0]))
    mx = edist.max_over_ranks(3.0 + rank, torch.device("cpu"))
    # result delivery to one rank (NCCL-gather fallback of ResultSink, here on gloo with host tensors)
    sink_ok = True
    if total % world == 0:
        sink = edist.ResultSink(tuple(local.shape), torch.float32, torch.device("cpu"), dst=0, slots=2)
        assert not sink.direct and sink.describe()["mode"].startswith("NCCL gather")
        for k in range(3):
            sink.before_write(k)
            sink.push(k, local + float(k))
        sink.drain()
        buf = sink.buffer()
        if rank == 0:
            sink_ok = torch.equal(buf[0].reshape(full.shape), full + 2.0) and torch.equal(buf[1].reshape(full.shape), full + 1.0)
        else:
            sink_ok = buf is None
    q.put((rank, torch.equal(got, full) and sink_ok, torch.equal(got2, full), acc.tolist(), mx))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7, 1])
def test_shard_gather_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, acc, mx in res:
        assert ok1 and ok2, rank
        assert acc == [3.0, 20.0]
        assert mx == 4.0


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 4, 8):
            spans = [edist.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
