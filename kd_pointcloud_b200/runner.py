"""Inference runner: the whole Bi-PointFlowNet forward captured ONCE in a CUDA graph.

The forward is ~350 small launches (kdpc kernels + cuBLAS GEMMs for the 1x1 convolutions); in
eager mode the Python/dispatcher overhead per launch exceeds most kernels' run time on a B200.
Shapes are static (N = 8192, levels hard-coded: models_bid_pointconv.py:31,40,49,58), so the
runner captures the forward + EPE3D into a graph over static input buffers and replays it.

``run_host`` is the public end-to-end call: pinned host batch in, host scalar (EPE3D) + device
flow out, with the host->device copies and the device->host read inside.
"""
from __future__ import annotations

import os
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


class PipelinedFlowRunner:
    """Steady-state inference with the sampling pyramid of batch i+1 overlapped with the forward of batch i.

    The FPS pyramid (4 levels, both clouds) needs only the input coordinates and is a latency chain: 1.3 ms of a
    5.4 ms step during which at most 64 SMs do anything.  Here a step is two CUDA graphs on two streams:

        stream A:  G_main[slot]   the forward of batch i consuming its pyramid        (persistent kernels capped to
                                                                                        ``num_sms - 2B`` CTAs)
        stream B:  G_fps[slot^1]  the pyramid of batch i+1 with the ONE-CTA-per-cloud FPS kernel (2B SMs), followed by
                                  every neighbour search that needs input coordinates only (14 of the forward's 22 kNN
                                  sets and 8 of its 11 spatial sorts); the forward finds them in the kNN cache

    with events in both directions (G_main(i) waits for G_fps(i); G_fps(i+2) waits for G_main(i), which reads the same
    pyramid buffers).  The persistent tcgen05 kernels assign tiles statically to their CTAs, so they must not be
    launched wider than the SMs the FPS CTAs leave free: ``kdpc_set_sm_limit`` caps their grids while the graphs are
    captured (work PLANS - split-K etc. - do not depend on the cap, so results are bit-identical to FlowRunner's).
    Every batch still gets the whole forward; only the order across batches changes.  Two slots of static buffers.
    """

    def __init__(self, model: torch.nn.Module, batch: int, npoints: int = 8192, device="cuda", precompute_knn: bool = True,
                 dual_forward: bool = False):
        self.model = model.eval()
        self.device = torch.device(device)
        self.batch = batch
        self.precompute_knn = precompute_knn and hasattr(model, "precompute_neighbours")
        self.neighbours = [None, None]
        self.static = [{k: torch.zeros(batch, npoints, 3, device=self.device) for k in KEYS} for _ in range(2)]
        self.geometry = [None, None]
        self.out_flow = [None, None]
        self.out_metrics = [None, None]
        self.g_fps = [None, None]
        self.g_main = [None, None]
        self.launches_per_step = 0
        self.stream_b = torch.cuda.Stream(device=self.device)
        self.stream_a = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
        self.dual_forward = dual_forward
        self.fps_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.main_done = [torch.cuda.Event(), torch.cuda.Event()]
        self._main_ran = [False, False]
        self.sm_limit = 0

    # ---- the two halves of a forward ----------------------------------------------------------------------------
    def _fps_part(self, slot: int):
        from . import _lib
        L = _lib.lib()
        s = self.static[slot]
        L.kdpc_fps_set_cluster(0)                           # one CTA per cloud: 2B SMs, leaves the rest to stream A
        try:
            with torch.no_grad():
                KF.clear_caches()
                self.geometry[slot] = self.model.sample_geometry(s["pos1"], s["pos2"])
                if self.precompute_knn:
                    # + every coordinate-only neighbour search of the forward (their results wait in the caches)
                    self.model.precompute_neighbours(self.geometry[slot])
                    self.neighbours[slot] = KF.snapshot_neighbour_caches()
                KF.clear_caches()
        finally:
            L.kdpc_fps_set_cluster(1)

    def _main_part(self, slot: int):
        KF.clear_caches()
        if self.neighbours[slot] is not None:
            KF.seed_neighbour_caches(self.neighbours[slot])
        s = self.static[slot]
        with torch.no_grad():
            flows = self.model(s["pos1"], s["pos2"], s["color1"], s["color2"], geometry=self.geometry[slot])[0]
            self.out_flow[slot] = flows[0]
            self.out_metrics[slot] = scene_flow_metrics(s["pos1"], flows[0], s["flow"])
        KF.clear_caches()

    def warmup_and_capture(self, sample: Dict[str, torch.Tensor], warmup: int = 2) -> bool:
        from . import _lib
        L = _lib.lib()
        dev = self.device
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        self.sm_limit = max(n_sm // 2, n_sm - 2 * self.batch)
        if os.environ.get("KDPC_SM_LIMIT"):                 # measurements only (tools/gpu_sweep_pipeline.sh)
            self.sm_limit = max(1, min(n_sm, int(os.environ["KDPC_SM_LIMIT"])))
        for slot in (0, 1):
            self.load(sample, slot)
        L.kdpc_set_sm_limit(self.sm_limit)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    for slot in (0, 1):
                        n0 = ops.LAUNCHES
                        self._fps_part(slot)
                        self._main_part(slot)
                        self.launches_per_step = ops.LAUNCHES - n0
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            # FOUR private memory pools, one per graph: a graph's temporaries are recycled inside its own pool only.
            # G_fps[s ^ 1] runs concurrently with G_main[s], and the OUTPUTS of G_fps[s] (the pyramid) must survive
            # a replay of G_fps[s ^ 1] - with a shared pool that replay's temporaries landed on them (seen as rare
            # 1e-5 .. 1e-2 EPE differences).
            for slot in (0, 1):
                gf = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gf):
                    self._fps_part(slot)
                gm = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gm):
                    self._main_part(slot)
                self.g_fps[slot], self.g_main[slot] = gf, gm
            self._keepalive = KF.weight_cache_tensors()
        finally:
            L.kdpc_set_sm_limit(0)
        torch.cuda.synchronize(dev)
        return True

    def load(self, batch: Dict[str, torch.Tensor], slot: int) -> int:
        nbytes = 0
        if self._main_ran[slot]:                            # the forward that last read this slot's inputs is done
            torch.cuda.current_stream(self.device).wait_event(self.main_done[slot])
        for k in KEYS:
            self.static[slot][k].copy_(batch[k], non_blocking=True)
            nbytes += batch[k].numel() * batch[k].element_size()
        return nbytes

    # ---- pipeline primitives --------------------------------------------------------------------------------------
    def launch_fps(self, slot: int, batch: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Sampling pyramid (+ coordinate-only kNN) of the batch in ``slot`` on stream B, after the previous forward that
        read this slot's inputs / pyramid.  ``batch``: device tensors copied into the slot's input buffers ON STREAM B
        first (so that the caller's stream never waits for a forward); without it the inputs written on the current
        stream (``load``) are used."""
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self.stream_b):
            self.stream_b.wait_event(ready)
            if self._main_ran[slot]:
                self.stream_b.wait_event(self.main_done[slot])
            if batch is not None:
                for k in KEYS:
                    self.static[slot][k].copy_(batch[k], non_blocking=True)
            self.g_fps[slot].replay()
            self.fps_done[slot].record(self.stream_b)

    def launch_main(self, slot: int) -> None:
        main = torch.cuda.current_stream(self.device)
        if self.dual_forward:
            # the two slots' forwards on their OWN streams: forward i+1 may start while forward i is still running; the
            # many-CTA kernels of one (kNN, sorts, interpolation) fill the issue slots the other's one-CTA-per-SM tcgen05
            # kernels leave idle, and a persistent kernel's tail overlaps the next one's ramp
            st = self.stream_a[slot]
            ready = torch.cuda.Event()
            ready.record(main)                             # this slot's inputs (and the L2 flush) were queued on `main`
            with torch.cuda.stream(st):
                st.wait_event(ready)
                st.wait_event(self.fps_done[slot])
                self.g_main[slot].replay()
                self.main_done[slot].record(st)
        else:
            main.wait_event(self.fps_done[slot])
            self.g_main[slot].replay()
            self.main_done[slot].record(main)
        self._main_ran[slot] = True

    def join(self) -> None:
        """Make the current stream wait for everything the pipeline has in flight."""
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self.stream_b)
        for st in self.stream_a:
            main.wait_stream(st)

    def run_resident(self, batches, on_step=None) -> list:
        """Device-resident batches through the two-stream pipeline; returns per-batch device metric tensors (clones).
        ``on_step(i)`` is called between steps on the main stream (bench.py flushes L2 there)."""
        out = []
        n = len(batches)
        if n == 0:
            return out
        self.launch_fps(0, batches[0])
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                self.launch_fps(slot ^ 1, batches[i + 1])  # overlaps launch_main(slot) below
            if on_step is not None:
                on_step(i)
            self.launch_main(slot)
            if self.dual_forward:
                with torch.cuda.stream(self.stream_a[slot]):
                    out.append(self.out_metrics[slot].clone())
            else:
                out.append(self.out_metrics[slot].clone())
        self.join()
        return out

    def run_host_pipelined(self, host_batches) -> list:
        """End to end over a stream of pinned host batches: H2D of batch i+1 (copy stream) and its sampling pyramid
        (stream B) overlap the forward of batch i; returns the EPE3D per batch."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._res_host = [torch.zeros(6, pin_memory=True) for _ in range(2)]
        copy = self._copy_stream
        h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
        res = []
        it = iter(host_batches)

        def upload(batch, slot):
            with torch.cuda.stream(copy):
                if self._main_ran[slot]:
                    copy.wait_event(self.main_done[slot])  # the forward that read this slot's inputs is done
                for k in KEYS:
                    self.static[slot][k].copy_(batch[k], non_blocking=True)
                h2d_done[slot].record(copy)

        def start_fps(slot):
            with torch.cuda.stream(self.stream_b):
                self.stream_b.wait_event(h2d_done[slot])
                if self._main_ran[slot]:
                    self.stream_b.wait_event(self.main_done[slot])
                self.g_fps[slot].replay()
                self.fps_done[slot].record(self.stream_b)

        nxt = next(it, None)
        if nxt is None:
            return res
        upload(nxt, 0)
        start_fps(0)
        i = 0
        pending = []
        while nxt is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, slot ^ 1)
                start_fps(slot ^ 1)
            if len(pending) == 2:
                j, ev = pending.pop(0)
                ev.synchronize()
                res.append(float(self._res_host[j][0]))
            main.wait_event(h2d_done[slot])
            self.launch_main(slot)
            rs = self.stream_a[slot] if self.dual_forward else main       # the stream this slot's forward runs on
            with torch.cuda.stream(rs):
                self._res_host[slot].copy_(self.out_metrics[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(rs)
            pending.append((slot, ev))
            i += 1
        for j, ev in pending:
            ev.synchronize()
            res.append(float(self._res_host[j][0]))
        self.join()
        return res
