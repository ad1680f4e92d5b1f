"""CPU tests (gloo, world_size 2) for the N > 1 host logic: batch sharding and the flat gradient
all-reduce must reproduce the single-process gradient over the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kd_pointcloud_b200.sharding import FlatGradAllReduce, broadcast_parameters, shard_batch, shard_range


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
    # numpy: pickled by value (tensors travel as shared-memory fds, which die with this process)
    q.put((rank, [None if p.grad is None else p.grad.numpy().copy() for p in m.parameters()], red.numel))
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
                assert torch.allclose(torch.from_numpy(a), b, rtol=1e-5, atol=1e-6), rank


def _worker_uneven(rank, world, port, q):
    """Uneven shards (5 + 3 of 8) with per-rank MEAN losses, ranks seeded differently: the broadcast makes the replicas
    identical, the batch-size weighting makes the result the gradient of the global-batch mean."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    a, b = (0, 5) if rank == 0 else (5, 8)
    mine = {k: v[a:b] for k, v in batch.items()}
    torch.manual_seed(100 + rank)                            # replicas start DIFFERENT ...
    m = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.BatchNorm1d(16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    m[1].running_mean.fill_(float(rank))
    red = FlatGradAllReduce(m.parameters(), module=m, local_batch=b - a)      # ... and are made equal here
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.eval()
    loss = ((m(mine["x"]) - mine["y"]) ** 2).sum(dim=1).mean()                # per-rank batch MEAN (loss_functions.py:22)
    loss.backward()
    red()
    # numpy: pickled by value (tensors travel as shared-memory fds, which die with this process)
    q.put((rank, {k: v.numpy() for k, v in sd.items()}, [p.grad.numpy().copy() for p in m.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_uneven_shards_give_the_global_batch_mean_gradient():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_uneven, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, sd0, g0), (_, sd1, g1) = res
    sd0, sd1 = ({k: torch.from_numpy(v) for k, v in sd.items()} for sd in (sd0, sd1))
    g0, g1 = ([torch.from_numpy(v) for v in gs] for gs in (g0, g1))
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k                # parameters AND buffers (running_mean) came from rank 0
    assert float(sd1["1.running_mean"][0]) == 0.0
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    m = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.BatchNorm1d(16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    m.load_state_dict(sd0)
    m.eval()
    ((m(batch["x"]) - batch["y"]) ** 2).sum(dim=1).mean().backward()
    for a, b, ref in zip(g0, g1, [p.grad for p in m.parameters()]):
        assert torch.allclose(a, ref, rtol=1e-5, atol=1e-6) and torch.equal(a, b)


def test_broadcast_parameters_is_a_noop_without_a_process_group():
    m = _model()
    before = [p.clone() for p in m.parameters()]
    broadcast_parameters(m)
    assert all(torch.equal(a, b) for a, b in zip(before, m.parameters()))


def _worker_sum_mode(rank, world, port, q):
    """A loss that mixes a batch MEAN with a batch SUM (the KD hint term, loss_functions.py:213-214): ranks scale their
    means by the global batch and ADD gradients (mode='sum')."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    a, b = (0, 5) if rank == 0 else (5, 8)
    m = _model()
    red = FlatGradAllReduce(m.parameters(), module=m, local_batch=b - a, mode="sum")
    out = m(batch["x"][a:b])
    loss = ((out - batch["y"][a:b]) ** 2).sum(dim=1).sum() / red.global_batch + 0.1 * (out ** 2).sum()
    loss.backward()
    red()
    q.put((rank, red.global_batch, [None if p.grad is None else p.grad.numpy().copy() for p in m.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_sum_mode_reproduces_a_loss_with_mean_and_sum_terms():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sum_mode, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    batch = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 3, generator=g)}
    m = _model()
    out = m(batch["x"])
    (((out - batch["y"]) ** 2).sum(dim=1).mean() + 0.1 * (out ** 2).sum()).backward()
    for rank, gb, grads in res:
        assert gb == 8
        for a, p in zip(grads, m.parameters()):
            assert (a is None) == (p.grad is None)
            if a is not None:
                assert torch.allclose(torch.from_numpy(a), p.grad, rtol=1e-5, atol=1e-6), rank
