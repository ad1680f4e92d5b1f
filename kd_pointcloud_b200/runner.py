"""Inference runner: the whole Bi-PointFlowNet forward captured ONCE in a CUDA graph.

The forward is ~350 small launches (kdpc kernels + cuBLAS GEMMs for the 1x1 convolutions); in
eager mode the Python/dispatcher overhead per launch exceeds most kernels' run time on a B200.
Shapes are static (N = 8192, levels hard-coded: models_bid_pointconv.py:31,40,49,58), so the
runner captures the forward + EPE3D into a graph over static input buffers and replays it.

``run_host`` is the public end-to-end call: pinned host batch in, host scalar (EPE3D) + device
flow out, with the host->device copies and the device->host read inside.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import functional as KF
from . import ops
from .evaluation_utils import scene_flow_metrics

KEYS = ("pos1", "pos2", "color1", "color2", "flow")


class FlowRunner:
    def __init__(self, model: torch.nn.Module, batch: int, npoints: int = 8192, device="cuda", use_graph: bool = True):
        self.model = model.eval()
        self.device = torch.device(device)
        self.static: Dict[str, torch.Tensor] = {k: torch.zeros(batch, npoints, 3, device=self.device) for k in KEYS}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out_flow: Optional[torch.Tensor] = None
        self.out_epe: Optional[torch.Tensor] = None
        self.out_metrics: Optional[torch.Tensor] = None    # EPE3D, Acc3DS, Acc3DR, Outliers3D, EPE2D, Acc2D (device)
        self.launches_per_step = 0
        self.use_graph = use_graph
        self._metrics_host = torch.zeros(6, pin_memory=True)

    def _forward(self):
        KF.clear_caches()                                  # never reuse indices across batches
        s = self.static
        with torch.no_grad():
            flows = self.model(s["pos1"], s["pos2"], s["color1"], s["color2"])[0]
            self.out_flow = flows[0]
            # evaluate_bid_pointconv.py:121-145 in one kernel (csrc/metrics.cu): EPE3D is element 0
            self.out_metrics = scene_flow_metrics(s["pos1"], flows[0], s["flow"])
            self.out_epe = self.out_metrics[:1]
        KF.clear_caches()

    def warmup_and_capture(self, sample: Dict[str, torch.Tensor], warmup: int = 2) -> bool:
        self.load(sample)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                n0 = ops.LAUNCHES
                self._forward()
                self.launches_per_step = ops.LAUNCHES - n0
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if not self.use_graph:
            return False
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._forward()
            self.graph = g
            # the graph bakes in the addresses of the cached packed weights / folded BN affines it hit: keep them alive
            # past LRU evictions and clear_caches(weights=True).  Weights re-loaded in place after capture are NOT seen
            # by the graph (no pack kernel is recorded for cache hits): call warmup_and_capture() again.
            self._keepalive = KF.weight_cache_tensors()
        except Exception as e:                             # eager still works; report it
            print(f"[kdpc] CUDA graph capture failed, running eagerly: {type(e).__name__}: {e}")
            self.graph = None
            torch.cuda.synchronize(self.device)
        return self.graph is not None

    def load(self, batch: Dict[str, torch.Tensor]) -> int:
        """Copy a (host or device) batch into the static buffers; returns bytes copied."""
        nbytes = 0
        for k in KEYS:
            self.static[k].copy_(batch[k], non_blocking=True)
            nbytes += batch[k].numel() * batch[k].element_size()
        return nbytes

    def step(self) -> None:
        """One forward over whatever is in the static buffers (device-resident path)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._forward()

    def run_host_pipelined(self, host_batches) -> list:
        """End-to-end over a STREAM of pinned host batches, software-pipelined: while the graph of batch i runs, batch
        i+1 travels host->device on a copy stream into one of two staging sets; a step then starts with a device-side
        copy staging -> static buffers (microseconds), replays the graph and reads the six metrics back.  Every step
        still pays its own H2D and D2H - they just no longer sit in front of the kernels.  Returns the EPE3D per batch."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = [{k: torch.empty_like(v) for k, v in self.static.items()} for _ in range(2)]
            self._res_host = [torch.zeros(6, pin_memory=True) for _ in range(2)]
        copy = self._copy_stream
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        res_done = [torch.cuda.Event(), torch.cuda.Event()]
        out, pending = [], []
        it = iter(host_batches)

        def upload(batch, slot, first_use):
            with torch.cuda.stream(copy):
                if not first_use:
                    copy.wait_event(consumed[slot])        # the step that used this staging set has copied it out
                for k in KEYS:
                    self._staging[slot][k].copy_(batch[k], non_blocking=True)
                h2d_done[slot].record(copy)

        nxt = next(it, None)
        if nxt is not None:
            upload(nxt, 0, True)
        i = 0
        while nxt is not None:
            slot = i & 1
            cur, nxt = nxt, next(it, None)
            if nxt is not None:
                upload(nxt, slot ^ 1, i == 0)              # overlaps this step's kernels
            main.wait_event(h2d_done[slot])
            for k in KEYS:
                self.static[k].copy_(self._staging[slot][k], non_blocking=True)
            consumed[slot].record(main)
            if len(pending) == 2:                          # the host slot about to be rewritten: collect its result first
                j, ev = pending.pop(0)
                ev.synchronize()
                out.append(float(self._res_host[j][0]))
            self.step()
            self._res_host[slot].copy_(self.out_metrics, non_blocking=True)
            res_done[slot] = torch.cuda.Event()
            res_done[slot].record(main)
            pending.append((slot, res_done[slot]))
            i += 1
        for j, ev in pending:
            ev.synchronize()
            out.append(float(self._res_host[j][0]))
        return out

    def run_host(self, host_batch: Dict[str, torch.Tensor]) -> float:
        """End-to-end: pinned host batch -> EPE3D as a Python float (H2D + forward + D2H)."""
        self.load(host_batch)
        self.step()
        self._metrics_host.copy_(self.out_metrics, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return float(self._metrics_host[0])
