"""The knowledge-distillation training step of distilTrain.py:156-185 as one function.

    teacher.eval(); with no_grad: teacher forward            (distilTrain.py:165-167)
    student.train(); student forward                         (:168-169; BatchNorm1d uses batch stats)
    loss = cross_biDirection_loss_ht(..., gamma=0.3, beta=0.8, layer=[2,3])   (:174)
    loss.backward(); [gradient all-reduce]; optimizer.step(); zero_grad       (:180-182)

The reference syncs the host twice per step (``loss.cpu()`` at :179 and :184); this returns the
device scalar and leaves the read to the caller.
"""
from __future__ import annotations

import copy
import os
from typing import Callable, Dict, Optional

import torch

from . import functional as KF
from . import losses


OVERLAP_TEACHER = os.environ.get("KDPC_OVERLAP_TEACHER", "1") != "0"
_SIDE = {}


def _teacher_and_student_forward(teacher, student, p1, p2, c1, c2):
    """Teacher forward (no grad) and student forward of one KD step.  Both see the same clouds, so the sampling pyramid and
    every coordinate-only neighbour set are computed ONCE up front; the teacher's forward then runs on a side stream while
    the student's forward is issued on the current one (the teacher is the inference path: one-CTA-per-SM tcgen05 kernels
    with idle issue slots; the student's training path is mostly many-CTA kernels).  Joined before the loss."""
    fused = (OVERLAP_TEACHER and p1.is_cuda and hasattr(teacher, "sample_geometry") and hasattr(teacher, "precompute_neighbours")
             and getattr(teacher, "level1", None) is not None and getattr(student, "level1", None) is not None
             and teacher.level1.npoint == student.level1.npoint)
    teacher.eval()
    if not fused:
        with torch.no_grad():
            t_out = teacher(p1, p2, c1, c2)
        student.train()
        return t_out, student(p1, p2, c1, c2)
    dev = p1.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE.get(dev)
    if side is None:
        side = _SIDE[dev] = torch.cuda.Stream(device=dev)
    with torch.no_grad():
        geo = teacher.sample_geometry(p1, p2)
        teacher.precompute_neighbours(geo)                 # shared, read-only from here on (functional's kNN / sort caches)
    side.wait_stream(main)
    with torch.cuda.stream(side), torch.no_grad():
        t_out = teacher(p1, p2, c1, c2, geometry=geo)
    student.train()
    s_out = student(p1, p2, c1, c2, geometry=geo)
    main.wait_stream(side)
    for group in t_out:                                     # produced on the side stream, consumed (loss) on this one
        for t in group:
            if isinstance(t, torch.Tensor):
                t.record_stream(main)
    return t_out, s_out


def _forward_backward(teacher, student, batch, optimizer, gamma, beta, layers, hint_mode, global_batch=None) -> torch.Tensor:
    KF.clear_caches()
    p1, p2, c1, c2, flow = batch["pos1"], batch["pos2"], batch["color1"], batch["color2"], batch["flow"]
    t_out, s_out = _teacher_and_student_forward(teacher, student, p1, p2, c1, c2)
    loss = losses.cross_biDirection_loss_ht(s_out[0], s_out[5], s_out[6], s_out[1], s_out[2], flow, t_out[0], t_out[5],
                                            t_out[6], t_out[1], t_out[2], gamma, beta, layer=layers, hint_mode=hint_mode,
                                            global_batch=global_batch)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    KF.clear_caches()
    return loss.detach()


def _global_batch(reducer) -> Optional[int]:
    """Batch-sharded training with a 'sum' reducer: batch means are taken over the global batch (see losses.py)."""
    return getattr(reducer, "global_batch", None) if getattr(reducer, "mode", None) == "sum" else None


def kd_step(teacher: torch.nn.Module, student: torch.nn.Module, batch: Dict[str, torch.Tensor],
            optimizer: torch.optim.Optimizer, reducer: Optional[Callable[[], None]] = None,
            gamma: float = 0.3, beta: float = 0.8, layers=(2, 3), hint_mode: str = "first") -> torch.Tensor:
    loss = _forward_backward(teacher, student, batch, optimizer, gamma, beta, layers, hint_mode, _global_batch(reducer))
    if reducer is not None:
        reducer()
    optimizer.step()
    return loss


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


class KdpcAdam(torch.optim.Optimizer):
    """torch.optim.Adam (L2 weight decay, bias correction, no amsgrad; distilTrain.py:134-135) as ONE kernel pass over all
    parameter tensors (csrc/adam.cu) instead of a dozen foreach launches over the 226 tensor lists: 1.9 -> ~0.1 ms per KD
    step.  Capturable by construction: the learning rate and the step count live in device tensors (``set_lr`` /
    schedulers that rewrite ``param_group['lr']`` in place keep working after a CUDA-graph capture), the pointer table
    travels in the kernel parameters.  State per parameter: ``exp_avg``, ``exp_avg_sq``; the step counter is one device
    scalar kept in the state of the group's first parameter (``kdpc_step``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        for g in self.param_groups:
            dev = g["params"][0].device
            if not isinstance(g["lr"], torch.Tensor):
                g["lr"] = torch.tensor(float(g["lr"]), device=dev)

    @torch.no_grad()
    def step(self, closure=None):
        import ctypes
        from . import _lib
        loss = closure() if closure is not None else None
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.grad is not None]
            if not ps:
                continue
            first = g["params"][0]
            if "kdpc_step" not in self.state[first]:
                self.state[first]["kdpc_step"] = torch.zeros((), dtype=torch.float32, device=first.device)
            step = self.state[first]["kdpc_step"]
            for p in ps:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                    raise ValueError("kdpc: KdpcAdam needs contiguous float32 parameters and gradients")
            n = len(ps)
            arr = ctypes.c_void_p * n
            lr = g["lr"] if isinstance(g["lr"], torch.Tensor) else torch.tensor(float(g["lr"]), device=first.device)
            with torch.cuda.device(first.device):
                rc = _lib.lib().kdpc_adam_step(
                    n, arr(*[p.data_ptr() for p in ps]), arr(*[p.grad.data_ptr() for p in ps]),
                    arr(*[self.state[p]["exp_avg"].data_ptr() for p in ps]), arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                    (ctypes.c_longlong * n)(*[p.numel() for p in ps]), ctypes.c_void_p(lr.data_ptr()), float(g["betas"][0]),
                    float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), ctypes.c_void_p(step.data_ptr()),
                    ctypes.c_void_p(torch.cuda.current_stream(first.device).cuda_stream))
            _lib.check(rc, "kdpc_adam_step")
            # the kernel wrote through raw pointers: move the tensors' version counters like an in-place torch op would, so
            # that everything keyed on (data_ptr, version) - packed weights, folded affines - is re-derived from the new
            # values (without this the next forward silently used the OLD packed student weights: the loss fell 391 ->
            # 379 -> 369 instead of 391 -> 339 -> 286; torch's own fused Adam showed the same symptom)
            for p in ps:
                torch.autograd.graph.increment_version(p)
        return loss


def make_capturable_adam(params, lr: float = 1e-3, **kw) -> torch.optim.Optimizer:
    """Adam for ``GraphedKDStep``: ``capturable=True`` and the learning rate as a DEVICE tensor, so that the reference's
    schedule (StepLR + the LEARNING_RATE_CLIP rewrite of ``param_group['lr']``, distilTrain.py:130-140) keeps working
    after capture: ``set_lr`` / an in-place scheduler update changes what the captured kernels read."""
    params = list(params)
    dev = params[0].device
    # Default: KdpcAdam (one kernel pass, csrc/adam.cu; KDPC_ADAM=torch selects torch.optim.Adam's capturable foreach path:
    # same updates, ~1.4 ms slower per step).  torch's `fused=True` variant (KDPC_ADAM_FUSED=1) trained visibly slower in
    # tools/cmp_adam.py - the same symptom KdpcAdam showed before it moved the parameters' version counters: the caches of
    # packed weights are keyed on (data_ptr, version) and kept serving the OLD student weights; and an explicit
    # fused=False silently selects the per-tensor loop (+4 ms), so neither flag is ever passed by default.
    import os
    if os.environ.get("KDPC_ADAM", "kdpc") == "kdpc" and not kw.get("amsgrad", False):
        return KdpcAdam(params, lr=lr, **{k: v for k, v in kw.items() if k in ("betas", "eps", "weight_decay")})
    if os.environ.get("KDPC_ADAM_FUSED", "0") == "1":
        kw.setdefault("fused", True)
    return torch.optim.Adam(params, lr=torch.tensor(float(lr), device=dev), capturable=True, **kw)


def set_lr(optimizer: torch.optim.Optimizer, lr: float) -> None:
    """``for g in optimizer.param_groups: g['lr'] = lr`` (distilTrain.py:139-140) that also reaches a captured graph."""
    for g in optimizer.param_groups:
        if isinstance(g["lr"], torch.Tensor):
            g["lr"].fill_(float(lr))
        else:
            g["lr"] = float(lr)


class GraphedKDStep:
    """``kd_step`` captured in CUDA graphs and replayed per batch: the eager step issues ~1600 kernel launches plus the
    autograd bookkeeping from Python and is bound by the host (71 ms of CPU per 73 ms of GPU time, tools/prof_train.py).

    * Without a gradient ``reducer``: ONE graph (teacher forward, student forward + backward, fused loss, Adam).
    * With a ``reducer`` (multi-GPU): TWO graphs — forward + backward into static ``.grad`` buffers, then the optimizer —
      with the NCCL all-reduce issued EAGERLY between them (capturing the collective hung the 2-GPU run in round 1; one
      31.8 MB all-reduce per step costs one launch from the host, the ~1600 other launches are replayed).

    Contract:
      * the optimizer must be ``capturable``; its lr is baked in unless it is a device tensor — use
        ``make_capturable_adam`` / ``set_lr`` so that LR schedules keep working after capture;
      * construction runs warm-up steps on ``example_batch``; student weights, BatchNorm statistics and optimizer state
        are RESTORED afterwards (``restore_after_warmup=True``), so building the stepper does not train;
      * ``step`` returns a CLONE of the static loss; a batch whose shapes differ from the captured ones (a short last
        batch) runs the eager step;
      * replays change weights without moving tensor versions: every replay bumps ``functional``'s weight epoch, so a
        later ``student.eval()`` forward re-derives packed weights / folded BN affines instead of using stale ones;
      * the graphs bake in the addresses of cached weight-derived tensors (teacher packs, folded affines): they are kept
        alive in ``self._keepalive``.  Weights must not be re-loaded in place after capture — call ``recapture()``.
    """

    def __init__(self, teacher: torch.nn.Module, student: torch.nn.Module, example_batch: Dict[str, torch.Tensor],
                 optimizer: torch.optim.Optimizer, reducer: Optional[Callable[[], None]] = None, warmup: int = 3,
                 restore_after_warmup: bool = True, **kw):
        self.teacher, self.student, self.optimizer, self.reducer = teacher, student, optimizer, reducer
        self.kw = {"gamma": kw.get("gamma", 0.3), "beta": kw.get("beta", 0.8), "layers": kw.get("layers", (2, 3)),
                   "hint_mode": kw.get("hint_mode", "first")}
        self.static = {k: v.clone() for k, v in example_batch.items()}
        self.loss: Optional[torch.Tensor] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None          # forward + backward (+ optimizer when no reducer)
        self.graph_opt: Optional[torch.cuda.CUDAGraph] = None      # optimizer (only with a reducer)
        self._keepalive = []
        self._warmup = max(3, warmup)
        self._restore = restore_after_warmup
        self.recapture()

    def _eager(self, batch) -> torch.Tensor:
        return kd_step(self.teacher, self.student, batch, self.optimizer, self.reducer, **self.kw)

    def recapture(self) -> None:
        """(Re)build the graphs from the CURRENT weights (call after load_state_dict or any in-place weight change)."""
        self.graph = self.graph_opt = None
        dev = next(self.student.parameters()).device
        saved = None
        if self._restore:
            saved = (copy.deepcopy(self.student.state_dict()), copy.deepcopy(self.optimizer.state_dict()))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):                   # allocator, caches and Adam state reach their steady state
                self._eager(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        try:
            g = torch.cuda.CUDAGraph()
            if self.reducer is None:
                with torch.cuda.graph(g):
                    self.loss = self._eager(self.static)
                self.graph = g
            else:
                with torch.cuda.graph(g):
                    self.loss = _forward_backward(self.teacher, self.student, self.static, self.optimizer,
                                                  global_batch=_global_batch(self.reducer), **self.kw)
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    self.optimizer.step()
                self.graph, self.graph_opt = g, g2
            self._keepalive = KF.weight_cache_tensors()
        except Exception as e:                             # eager still works; report it
            print(f"[kdpc] CUDA graph capture of the KD step failed, running eagerly: {type(e).__name__}: {e}")
            self.graph = self.graph_opt = None
            torch.cuda.synchronize(dev)
        if saved is not None:                              # building the stepper must not train the student
            with torch.no_grad():
                self.student.load_state_dict(saved[0])     # in place: the captured addresses stay valid
                cur = self.optimizer.state_dict()
                for k, st in saved[1]["state"].items():
                    for name, v in st.items():
                        if isinstance(v, torch.Tensor) and k in cur["state"]:
                            cur["state"][k][name].copy_(v)
                missing = [k for k in cur["state"] if k not in saved[1]["state"]]
                for k in missing:                          # optimizer state created by the warm-up: back to its initial value
                    for name, v in cur["state"][k].items():
                        if isinstance(v, torch.Tensor):
                            v.zero_()
            KF.bump_weight_epoch()

    def _shapes_match(self, batch) -> bool:
        return all(k in batch and batch[k].shape == v.shape for k, v in self.static.items())

    def step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.graph is None or not self._shapes_match(batch):
            params = [p for g in self.optimizer.param_groups for p in g["params"]]
            static_grads = [p.grad for p in params]        # the graphs write / read THESE buffers: keep them bound
            loss = self._eager(batch)
            if self.graph is not None:
                for p, g in zip(params, static_grads):
                    p.grad = g
            KF.bump_weight_epoch()
            return loss
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        if self.graph_opt is not None:
            self.reducer()                                 # eager NCCL all-reduce over the static .grad buffers
            self.graph_opt.replay()
        KF.bump_weight_epoch()                             # weights moved on the device; tensor versions did not
        return self.loss.clone()
