"""Multi-GPU plumbing: one process per GPU, scene pairs sharded by batch.

Inference needs NO collective (every kernel is per batch element).  Training adds exactly
one: a sum all-reduce of the gradients.  The reference's only multi-GPU mechanism is single-process
``nn.DataParallel`` (distilTrain.py:108-114: replicate weights + scatter + gather + reduce on GPU0
every step); here each rank owns its shard end to end and the gradients travel once, as ONE flat
fp32 buffer (7 958 604 elements = 31.8 MB for the student) over NCCL / NVLink.

80 of the 306 parameter tensors never receive a gradient (CrossLayerLight.bias1/bias2,
WeightNet.mlp_bns.*) — DistributedDataParallel would need find_unused_parameters; the
flat reducer simply skips ``grad is None`` (the set is structural, hence identical on all ranks).
BatchNorm1d statistics stay per replica, as with the reference's DataParallel.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``total`` independent items; the first ``total % world`` ranks
    get one extra."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    n = next(iter(batch.values())).shape[0]
    a, b = shard_range(n, rank, world)
    return {k: v[a:b] for k, v in batch.items()}


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make every rank start from rank ``src``'s parameters AND buffers (BatchNorm statistics): replicas that load
    different checkpoints, or a ``--pretrain`` given to one rank only, would otherwise diverge silently.  One flat
    broadcast per dtype."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        torch._foreach_copy_(ts, [c.view_as(t) for c, t in zip(flat.split([t.numel() for t in ts]), ts)])


class FlatGradAllReduce:
    """Average gradients across ranks with one all-reduce over a persistent flat buffer.

    ``module``: when given, its parameters and buffers are broadcast from rank 0 at construction (see
    ``broadcast_parameters``).  ``local_batch``: the number of pairs this rank contributes per step; the result is the
    GLOBAL-batch mean ``sum_r n_r g_r / sum_r n_r`` (for equal shards this is the plain average; a rank may pass 0 and
    contribute nothing).  Losses here are per-rank batch means (loss_functions.py:22 ``.mean()`` over B), which is what
    makes this weighting the concatenated-batch gradient."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 module: Optional[torch.nn.Module] = None, local_batch: Optional[int] = None, mode: str = "mean"):
        """mode='mean': per-rank losses are batch means; the result is the global-batch mean gradient.
        mode='sum' : per-rank losses are already scaled for the global batch (``self.global_batch`` is what their means
                     must divide by; losses with per-sample SUM terms, e.g. the KD hint loss, need this): gradients are
                     added, nothing is divided."""
        if mode not in ("mean", "sum"):
            raise ValueError(mode)
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.mode = mode
        self.global_batch: Optional[int] = local_batch
        self.flat: Optional[torch.Tensor] = None
        self.active: Optional[List[int]] = None
        self.weight = 1.0                                  # this rank's share of the global batch times world
        if module is not None:
            broadcast_parameters(module, 0, group)
        if local_batch is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
            dev = self.params[0].device if dist.get_backend(group) == "nccl" else "cpu"
            n = torch.tensor([float(local_batch)], dtype=torch.float64, device=dev)
            total = n.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
            if total.item() <= 0:
                raise RuntimeError("FlatGradAllReduce: the global batch is empty")
            self.global_batch = int(round(total.item()))
            if mode == "mean":
                self.weight = float(local_batch) * dist.get_world_size(group) / total.item()

    def _plan(self):
        self.active = [i for i, p in enumerate(self.params) if p.grad is not None]
        n = sum(self.params[i].numel() for i in self.active)
        ref = self.params[self.active[0]]
        self.flat = torch.zeros(n, dtype=torch.float32, device=ref.device)
        if dist.is_initialized():
            # the participating set is structural; verify once that all ranks agree on its size
            sizes = torch.tensor([n, len(self.active)], dtype=torch.int64, device=ref.device)
            lo, hi = sizes.clone(), sizes.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
            if not torch.equal(lo, hi):
                raise RuntimeError("FlatGradAllReduce: ranks disagree on which parameters have gradients")

    @property
    def numel(self) -> int:
        return 0 if self.flat is None else self.flat.numel()

    def __call__(self) -> None:
        if self.active is None:
            self._plan()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return
        grads = [self.params[i].grad for i in self.active]
        if any(g is None for g in grads):
            raise RuntimeError("FlatGradAllReduce: a parameter lost its gradient after planning")
        torch._foreach_copy_(list(self.flat.split([g.numel() for g in grads])), [g.reshape(-1) for g in grads])
        if self.weight != 1.0:
            self.flat.mul_(self.weight)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.mode == "mean":
            self.flat.div_(world)
        torch._foreach_copy_([g.view(-1) if g.is_contiguous() else g for g in grads],
                             [c.view_as(g) if not g.is_contiguous() else c for c, g in
                              zip(self.flat.split([g.numel() for g in grads]), grads)])
