"""CPU tests (gloo, world_size 2) for the N > 1 host logic: batch sharding and the flat gradient
all-reduce must reproduce the single-process gradient over the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kd_pointcloud_b200.sharding import FlatGradAllReduce, shard_batch, shard_range


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _model():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    m.unused = torch.nn.Parameter(torch.zeros(4))          # never gets a gradient (like CrossLayerLight.bias1)
    return m


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    mine = shard_batch(batch, rank, world)
    m = _model()
    red = FlatGradAllReduce(m.parameters())
    for _ in range(2):                                       # twice: the flat buffer is reused
        m.zero_grad(set_to_none=True)
        loss = ((m(mine["x"]) - mine["y"]) ** 2).sum() / 8 * world      # per-rank mean-of-global * world -> average = global
        loss.backward()
        red()
    q.put((rank, [None if p.grad is None else p.grad.clone() for p in m.parameters()], red.numel))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    m = _model()
    (((m(batch["x"]) - batch["y"]) ** 2).sum() / 8).backward()
    ref = [None if p.grad is None else p.grad for p in m.parameters()]
    for rank, grads, numel in res:
        assert numel == sum(r.numel() for r in ref if r is not None)
        for a, b in zip(grads, ref):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), rank
