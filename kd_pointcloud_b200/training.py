"""The knowledge-distillation training step of distilTrain.py:156-185 as one function.

    teacher.eval(); with no_grad: teacher forward            (distilTrain.py:165-167)
    student.train(); student forward                         (:168-169; BatchNorm1d uses batch stats)
    loss = cross_biDirection_loss_ht(..., gamma=0.3, beta=0.8, layer=[2,3])   (:174)
    loss.backward(); [gradient all-reduce]; optimizer.step(); zero_grad       (:180-182)

The reference syncs the host twice per step (``loss.cpu()`` at :179 and :184); this returns the
device scalar and leaves the read to the caller.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import functional as KF
from . import losses


def kd_step(teacher: torch.nn.Module, student: torch.nn.Module, batch: Dict[str, torch.Tensor],
            optimizer: torch.optim.Optimizer, reducer: Optional[Callable[[], None]] = None,
            gamma: float = 0.3, beta: float = 0.8, layers=(2, 3), hint_mode: str = "first") -> torch.Tensor:
    KF.clear_caches()
    p1, p2, c1, c2, flow = batch["pos1"], batch["pos2"], batch["color1"], batch["color2"], batch["flow"]
    teacher.eval()
    with torch.no_grad():
        t_out = teacher(p1, p2, c1, c2)
    student.train()
    s_out = student(p1, p2, c1, c2)
    loss = losses.cross_biDirection_loss_ht(s_out[0], s_out[5], s_out[6], s_out[1], s_out[2], flow, t_out[0], t_out[5],
                                            t_out[6], t_out[1], t_out[2], gamma, beta, layer=layers, hint_mode=hint_mode)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if reducer is not None:
        reducer()
    optimizer.step()
    KF.clear_caches()
    return loss.detach()


def supervised_step(model: torch.nn.Module, batch: Dict[str, torch.Tensor], optimizer: torch.optim.Optimizer,
                    reducer: Optional[Callable[[], None]] = None) -> torch.Tensor:
    """train_bid_pointconv.py:140-160: multiScaleLoss(pred_flows, flow, fps_pc1_idxs) + Adam."""
    KF.clear_caches()
    model.train()
    out = model(batch["pos1"], batch["pos2"], batch["color1"], batch["color2"])
    loss = losses.multiScaleLoss(out[0], batch["flow"], out[1])
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if reducer is not None:
        reducer()
    optimizer.step()
    KF.clear_caches()
    return loss.detach()


class GraphedKDStep:
    """``kd_step`` captured once in a CUDA graph (teacher forward, student forward + backward, fused loss, optional
    gradient all-reduce, Adam) and replayed per batch: the eager step issues ~1600 kernel launches plus the autograd
    bookkeeping from Python and is bound by the host (71 ms of CPU per 73 ms of GPU time, tools/prof_train.py).

    The optimizer must be created with ``capturable=True``; batches are copied into static input buffers.
    Falls back to the eager step (``self.graph is None``) when the capture fails, and ALWAYS runs eagerly when a gradient
    ``reducer`` is given: capturing the NCCL all-reduce hung the 2-GPU run (round 1), so multi-GPU training stays eager
    until that is understood."""

    def __init__(self, teacher: torch.nn.Module, student: torch.nn.Module, example_batch: Dict[str, torch.Tensor],
                 optimizer: torch.optim.Optimizer, reducer: Optional[Callable[[], None]] = None, warmup: int = 3, **kw):
        self.teacher, self.student, self.optimizer, self.reducer, self.kw = teacher, student, optimizer, reducer, kw
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.loss: Optional[torch.Tensor] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        if reducer is not None:
            return
        dev = next(student.parameters()).device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(3, warmup)):                 # allocator, caches and Adam state reach their steady state
                kd_step(teacher, student, self.static, optimizer, reducer, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.loss = kd_step(teacher, student, self.static, optimizer, reducer, **kw)
            self.graph = g
        except Exception as e:                             # eager still works; report it
            print(f"[kdpc] CUDA graph capture of the KD step failed, running eagerly: {type(e).__name__}: {e}")
            torch.cuda.synchronize(dev)

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.graph is None:
            return kd_step(self.teacher, self.student, batch, self.optimizer, self.reducer, **self.kw)
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.loss
