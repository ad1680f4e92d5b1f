#!/usr/bin/env python
"""Headline benchmark: scene-flow pairs/s @ 8192 points (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU arm (oracle port)

Workload (configs[2]): Bi-PointFlowNet (teacher, models_bid_pointconv.PointConvBidirection) eval
forward + EPE3D on FlyingThings3D-shaped synthetic 8192-point pairs, B = 8 per GPU, seeded
synthetic weights.  One step = one forward over one batch.  N > 1: pairs are independent, so the
batch is sharded over ranks with NO data-path collective (weak scaling); the only collective is
the timing reduction.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scene-flow pairs/sec @8192 pts"
UNIT = "pairs/s"
NPOINTS = 8192
MODEL_SEED = 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kdpc", choices=["kdpc", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="pairs per GPU per step")
    ap.add_argument("--workload", default="infer", choices=["infer", "kd_train"],
                    help="infer = configs[2] (the headline); kd_train = configs[3]/[4]: teacher fwd + student fwd/bwd + "
                         "fused KD loss + Adam, gradients all-reduced over NCCL when N > 1")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="time the single-graph runner (one batch at a time) instead of the two-stream pipeline that "
                         "overlaps the next batch's sampling pyramid with the current forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the GPU-side baseline leg (the unmodified reference model on its own torch layers + its own "
                         "sm_100a kernels, baseline/_ref + oracle/_ref)")
    ap.add_argument("--profile-one", action="store_true",
                    help="run ONE eager forward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    return ap.parse_args()


def workload_config(batch: int):
    """The ``config`` object: identical for the kdpc arm and the reference arm (same workload, same batch per step)."""
    return {"workload": "configs[2]: Bi-PointFlowNet teacher eval forward + EPE3D, FlyingThings3D-shaped synthetic 8192-pt "
                        "pairs, seeded synthetic weights", "npoints": NPOINTS, "pairs_per_device_per_step": batch,
            "batch_seed": "make_pairs(batch, 8192, seed=1234 + 1000*rank + step % 4)", "model_seed": MODEL_SEED}


def ncu_capture_of_roofline_kernel():
    """dram bytes / tensor-pipe share of the roofline kernel from the COMMITTED ncu capture of this build
    (profiles/rNN_ncu_pointconv.txt, written by tools/ncu_summary.py): newest round first."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_pointconv.txt")), reverse=True):
        vals = {}
        with open(path) as f:
            for line in f:
                m = re.match(r"(dram read|dram write|tensor pipe cycles active %|duration)\s+([0-9.,]+)\s*(\S*)", line)
                if m and m.group(1) not in vals:
                    v = float(m.group(2).replace(",", ""))
                    unit = m.group(3).lower()
                    mult = {"mbyte": 1e6, "gbyte": 1e9, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
                    vals[m.group(1)] = v * mult
        if "dram read" in vals and "dram write" in vals:
            return {"traffic": vals["dram read"] + vals["dram write"], "tensor_pipe_pct": vals.get("tensor pipe cycles active %"),
                    "source": os.path.relpath(path, ROOT)}
    return {"traffic": None, "tensor_pipe_pct": None, "source": None}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1600.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1600.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout: float = 5.0):
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout:
            time.sleep(0.05)

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                mx = float(parts[1])
                if t0 - 0.1 <= t <= t1 + 0.1:
                    sm.append(float(parts[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:                                       # timed region shorter than the sampling period
            for t, line in self.rows[-3:]:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- reference arm
def cpu_reference_forward(steps: int, warmup: int, seed0: int = 1234, batch: int = 1, pairs=None):
    """The reference's own algorithm on the host CPU: oracle/layers_ref.py (a restatement of
    pointconv_util.py / models_bid_pointconv.py pinned against the unmodified reference by
    tests/make_golden.py; kNN = matmul expansion + topk exactly as the reference does it;
    FPS/gather/group from oracle/kdpc_oracle.c because the reference has no CPU version).
    One step = ``batch`` pairs at 8192 points drawn exactly like the kdpc arm draws them; ``pairs`` (a dict of host
    tensors) replaces the drawn batch (used to evaluate the oracle on a pair the GPU arm has just processed)."""
    import torch
    from oracle import layers_ref as O
    from kd_pointcloud_b200.flownet import PointConvBidirection
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic_state_dict(PointConvBidirection().state_dict(), MODEL_SEED)
    times = []
    epe = None
    with torch.no_grad():
        for i in range(warmup + steps):
            d = pairs if pairs is not None else make_pairs(batch, NPOINTS, seed=seed0 + i % 4)
            t0 = time.perf_counter()
            flows = O.bid_pointconv_forward(sd, d["pos1"], d["pos2"], d["color1"], d["color2"], knn_impl="torch")[0]
            epe = torch.norm(flows[0].permute(0, 2, 1) - d["flow"], dim=2).mean().item()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / max(len(times), 1)
    nb = (pairs["pos1"].shape[0] if pairs is not None else batch)
    return nb / sec, sec, torch.get_num_threads(), epe


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, cores, _ = cpu_reference_forward(args.steps, args.warmup, batch=args.batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch),
        "details": {"host": "CPU only (one process on rank 0, all host threads)", "global_pairs_per_step": args.batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.batch} pairs (the kdpc arm's batch) per step; oracle/layers_ref.py with torch "
                                   "matmul+topk kNN exactly as the reference"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------- our arm
def kernel_rooflines(torch, dev, B, hbm_peak, tc_peak):
    """Live CUDA-event timing (L2 flushed before every launch) of the hot kernels at their largest model shapes.
    Algorithmic bytes / flops per SURVEY 8(d) / DESIGN.md section 4."""
    from kd_pointcloud_b200 import functional as KF
    from kd_pointcloud_b200 import pointconv_util as P
    from kd_pointcloud_b200.synth import make_pairs
    K = torch.ops.kdpc
    d = make_pairs(B, NPOINTS, seed=99, device=dev)
    xyz, xyz2 = d["pos1"], d["pos2"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=8):
        fn()
        ts = []
        for _ in range(iters):
            torch.cuda._sleep(400000)                      # the host runs ahead of the GPU: events bracket GPU execution only
            flush.zero_()                                  # evict L2 (126 MB) between timed launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2] * 1e-3

    out = {}
    N = NPOINTS
    # ---- fused PointConv, flow0 shape: [B,8192] points, K=9 neighbours, D=128 -> 128 (pointconv_util.py:231-258)
    D, Cout, Kn = 128, 128, 9
    idx9 = K.knn(xyz, xyz, Kn)
    feats = torch.randn(B, N, D, device=dev)
    wn = P.WeightNet(3, 16).to(dev)
    lin = torch.nn.Linear(16 * (D + 3), Cout).to(dev)
    wp = K.pack_weight(lin.weight.detach(), 1, D, 16)
    params = KF._weightnet_host_params(wn.mlp_convs)
    bias = lin.bias.detach()
    KF._knn_compute(Kn, xyz, xyz)                          # the model's kNN leaves the cloud's Morton order in the sort cache
    order = KF.morton_order(xyz)                           # (what PointConv.forward passes: functional.fused_pointconv)
    t = timeit(lambda: K.pointconv_fused(xyz, xyz, feats, idx9, params, wp, Cout, None, bias, 0.1, order))
    S = B * N
    flops = 2.0 * S * (D + 3) * Kn * 16 + 2.0 * S * 16 * (D + 3) * Cout + 2.0 * S * Kn * 216
    byts = B * (4 * N * Kn + 24 * N + 4 * N * D + 4 * N * Cout)
    out["pointconv_fused"] = {"shape": f"B={B} S=N={N} K={Kn} D={D}->{Cout}", "flops": flops, "bytes": byts, "sec": t,
                              "tflops": flops / t / 1e12, "frac_tensor": flops / t / 1e12 / tc_peak,
                              "gbs": byts / t / 1e9, "note": "fp32 result from 3 bf16 MMAs per product: tensor-pipe work is 3x the algorithmic flops of the Linear"}
    # ---- exact kNN (Morton sort of both clouds + best-first search), l0 cross-frame shape
    for kk in (32, 16, 9, 3):
        t = timeit(lambda: K.knn(xyz2, xyz, kk), iters=5)
        algk = B * (12 * (N + N) + 4 * N * kk)
        out[f"knn_k{kk}"] = {"shape": f"B={B} S=N={N} K={kk}", "bytes": algk, "sec": t, "gbs": algk / t / 1e9,
                             "gpairs_per_s": B * N * N / t / 1e9}
    # ---- fused cost volume, cross0 shape (pointconv_util.py:1826-1850)
    idx32 = K.knn(xyz2, xyz, 32)
    Dc = 32
    p1, p2 = torch.randn(B, N, Dc, device=dev), torch.randn(B, N, Dc, device=dev)
    pw, pb = torch.randn(Dc, 3, device=dev), torch.randn(Dc, device=dev)
    wpc = K.pack_weight(torch.randn(Dc, Dc, device=dev), 0, 0, 0)
    t = timeit(lambda: K.costvol_fused(xyz2, xyz, p1, p2, idx32, pw, pb, 0.1, wpc, Dc, pb, 0.1))
    algc = B * (4 * N * 32 + 24 * N + 2 * 4 * N * Dc + 4 * N * Dc)
    out["costvol_fused"] = {"shape": f"B={B} N={N} K=32 D={Dc}", "bytes": algc, "sec": t, "gbs": algc / t / 1e9,
                            "flops": 2.0 * B * N * 32 * Dc * (Dc + 3)}
    # ---- streaming 1x1 convolution on tcgen05 (flow0 mlp: 65536 x 128 -> 128)
    x = torch.randn(B * N, 128, device=dev)
    wpl = K.pack_weight(torch.randn(128, 128, device=dev), 0, 0, 0)
    sh = torch.randn(128, device=dev)
    t = timeit(lambda: K.linear_tc(x, wpl, 128, None, sh, 0.1, 1.0, 0.0, None))
    algl = B * N * (128 + 128) * 4
    out["linear_tc"] = {"shape": f"M={B * N} K=128 N=128", "bytes": algl, "sec": t, "gbs": algl / t / 1e9}
    # ---- grouping (training path; inference fuses it into PointConv)
    t = timeit(lambda: K.group_concat(xyz, xyz, feats, idx9))
    alg = B * (4 * N * Kn + 12 * (N + N) + 4 * N * D + 4 * N * Kn * (D + 3))
    out["group_concat"] = {"shape": f"B={B} S=N={N} K={Kn} D={D}", "bytes": alg, "sec": t, "gbs": alg / t / 1e9}
    # ---- point-major row gather (index_points_group, pointconv_util.py:122-133) and 3-NN interpolation (UpsampleFlow)
    f64 = torch.randn(B, N, 64, device=dev)
    idx16 = K.knn(xyz, xyz, 16)
    t = timeit(lambda: K.gather_rows(f64, idx16))
    alg = B * (4 * N * 16 + 4 * N * 64 + 4 * N * 16 * 64)
    out["gather_rows"] = {"shape": f"B={B} N={N} K=16 C=64", "bytes": alg, "sec": t, "gbs": alg / t / 1e9}
    fps_i = K.fps(xyz, 2048)
    sparse = K.gather_rows(xyz, fps_i)
    idx3 = K.knn(xyz, sparse, 3)
    fs = torch.randn(B, 2048, 64, device=dev)
    t = timeit(lambda: K.interp3(xyz, sparse, idx3, fs))
    alg = B * (4 * 2048 * 64 + 36 * N + 12 * 2048 + 4 * N * 64)
    out["interp3"] = {"shape": f"B={B} N={N} S=2048 C=64", "bytes": alg, "sec": t, "gbs": alg / t / 1e9}
    # ---- FPS: latency-bound (sequential arg-max)
    t = timeit(lambda: K.fps(xyz, 2048), iters=3)
    out["fps_8192_2048"] = {"shape": f"B={B} 8192->2048", "sec": t, "us_per_iter": t / 2047 * 1e6,
                            "bytes": B * (12 * N + 4 * 2048), "gbs": B * (12 * N + 4 * 2048) / t / 1e9}
    # ---- training path (configs[3]/[4]): aggregation backward, weight gradient, arg-max cost-volume backward
    grouped = K.group_concat(xyz, xyz, feats, idx9)
    wn_out = torch.rand(B, N, Kn, 16, device=dev)
    gagg = torch.randn(B, N, (D + 3) * 16, device=dev)
    t = timeit(lambda: K.pointconv_agg_grad(grouped, wn_out, gagg, True, True), iters=4)
    alg = 4 * B * N * (2 * Kn * (D + 3) + 2 * Kn * 16 + 16 * (D + 3))
    out["pointconv_agg_grad"] = {"shape": f"rows={B * N} K={Kn} C={D + 3}", "bytes": alg, "sec": t, "gbs": alg / t / 1e9}
    del grouped, wn_out, gagg
    dy, xr = torch.randn(B * N * 32, 32, device=dev), torch.randn(B * N * 32, 32, device=dev)
    t = timeit(lambda: K.linear_dw(dy, xr, True), iters=4)
    alg = 4 * B * N * 32 * 64
    out["linear_dw"] = {"shape": f"M={B * N * 32} N=32 K=32", "bytes": alg, "sec": t, "gbs": alg / t / 1e9}
    del dy, xr
    w32 = torch.randn(Dc, Dc, device=dev) / Dc ** 0.5
    gcv = torch.randn(B, N, Dc, device=dev)
    t = timeit(lambda: K.costvol_grad(p1, p2, idx32, w32, pb, 0.1, 0.1, gcv), iters=4)
    alg = 4 * B * N * (4 * Dc + 32 + 32 * Dc)
    out["costvol_grad"] = {"shape": f"B={B} N={N} K=32 D={Dc}", "bytes": alg, "sec": t, "gbs": alg / t / 1e9,
                           "note": "recomputing arg-max backward: FMA / latency bound, not HBM"}
    for k, v in out.items():
        v["frac_hbm"] = v["gbs"] / hbm_peak
    return out


def reference_gpu_forward(torch, dev, host_batch, B):
    """GPU-side baseline (reported, never on the product path): the UNMODIFIED models_bid_pointconv.py from baseline/_ref
    on the reference's own pointconv_util.py (torch eager) and its own CUDA kernels recompiled for sm_100a (oracle/_ref),
    same weights, same batch, same GPU.  1 warm-up + 2 timed forwards, CUDA events."""
    try:
        from oracle import ref_gpu
        if not ref_gpu.stock_available():
            return {"unavailable": "baseline/_ref or oracle/_ref not present (tools/install_reference.sh, oracle/Makefile)"}
        from kd_pointcloud_b200.synth import synthetic_state_dict
        mods = ref_gpu.load("stock")
        m = mods["models_bid_pointconv"].PointConvBidirection()
        m.load_state_dict(synthetic_state_dict(m.state_dict(), MODEL_SEED))
        m = m.to(dev).eval()
        d = {k: v.to(dev) for k, v in host_batch.items()}
        ts = []
        with torch.no_grad():
            for i in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                flows = m(d["pos1"], d["pos2"], d["color1"], d["color2"])[0]
                b.record()
                b.synchronize()
                if i:
                    ts.append(a.elapsed_time(b))
        epe = float(torch.norm(flows[0].permute(0, 2, 1) - d["flow"], dim=2).mean().item())
        ms = sum(ts) / len(ts)
        del m
        torch.cuda.empty_cache()
        return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "epe3d": epe,
                "what": "unmodified reference model + layer library (torch eager: matmul+topk kNN, cuDNN/cuBLAS fp32) + its own "
                        "pointnet2 kernels compiled for sm_100a; eager, device-resident inputs"}
    except Exception as e:                                  # a reported baseline must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_kdpc(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=kdpc) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False          # fp32 parity with the reference
    torch.backends.cudnn.allow_tf32 = False

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from kd_pointcloud_b200 import ops
    from kd_pointcloud_b200.flownet import PointConvBidirection
    from kd_pointcloud_b200.runner import FlowRunner, KEYS
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict

    sampler = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi takes ~1 s to emit
    B = args.batch
    model = PointConvBidirection()
    model.load_state_dict(synthetic_state_dict(model.state_dict(), MODEL_SEED))
    model = model.to(dev).eval()

    # a pool of DIFFERENT batches (per rank, per step) so nothing can be cached across steps
    pool = 4
    host = [{k: v.pin_memory() for k, v in make_pairs(B, NPOINTS, seed=1234 + 1000 * rank + i).items()} for i in range(pool)]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]

    if args.profile_one and not args.no_pipeline:
        # the FORWARD part of the pipelined step (stream A's graph), eagerly, between cudaProfilerStart/Stop
        from kd_pointcloud_b200 import _lib
        from kd_pointcloud_b200.runner import PipelinedFlowRunner
        pipe = PipelinedFlowRunner(model, B, NPOINTS, dev)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        pipe.sm_limit = max(n_sm // 2, n_sm - 2 * B)
        _lib.lib().kdpc_set_sm_limit(pipe.sm_limit)
        for slot in (0, 1):
            pipe.load(resident[slot], slot)
        for _ in range(2):
            pipe._fps_part(0)
            pipe._main_part(0)
        pipe._fps_part(1)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        pipe._main_part(1)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_one": "pipelined forward part", "epe3d": float(pipe.out_metrics[1][0].item())}))
        return
    if args.profile_one:
        runner = FlowRunner(model, B, NPOINTS, dev, use_graph=False)
        runner.warmup_and_capture(resident[0], warmup=2)
        runner.load(resident[1])
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        runner.step()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_one": True, "kdpc_calls": runner.launches_per_step, "epe3d": float(runner.out_epe.item())}))
        return

    runner = FlowRunner(model, B, NPOINTS, dev, use_graph=not args.no_graph)
    graphed = runner.warmup_and_capture(resident[0], warmup=2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pipelined = graphed and not args.no_pipeline and 2 * B <= 32
    pipe = None
    if pipelined:
        from kd_pointcloud_b200.runner import PipelinedFlowRunner
        pipe = PipelinedFlowRunner(model, B, NPOINTS, dev, dual_forward=os.environ.get("KDPC_DUAL_FORWARD", "1") == "1")
        pipe.warmup_and_capture(resident[0], warmup=1)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (value) -------------------------------------
    for i in range(args.warmup):
        runner.load(resident[i % pool])
        runner.step()
    barrier()
    if sampler:
        sampler.wait_first()
    t_wall0 = time.time()
    evs = []
    launches0 = ops.LAUNCHES
    for i in range(args.steps):
        runner.load(resident[i % pool])                    # device->device, outside the timed events
        flush.zero_()                                      # evict L2 between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        runner.step()
        b.record()
        evs.append((a, b))
    barrier()
    t_wall1 = time.time()
    t_clock0 = t_wall0                                       # clocks are sampled over BOTH timed regions (single graph, pipeline)
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    eager_launches = ops.LAUNCHES - launches0
    launches = runner.launches_per_step * args.steps if graphed else eager_launches
    epe_last = float(runner.out_epe.item())
    single_ms_step = dev_ms / args.steps

    # ---------------- the same K batches through the two-stream pipeline (the reported value) ----------------------
    # step i = forward of batch i (stream A) beside the sampling pyramid of batch i+1 (stream B).  The timed region starts
    # with batch 0's pyramid already computed and ends when batch K's pyramid is: K pyramids + K forwards = K whole
    # batches.  The L2 flush and the device-to-device load of the next batch are INSIDE the timed region.
    if pipe is not None:
        main = torch.cuda.current_stream(dev)
        for rep in range(2):                                 # rep 0: warm-up of the pipeline itself
            pipe.launch_fps(0, resident[0])
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_wall0 = time.time()
            a.record()
            for i in range(args.steps):
                slot = i & 1
                pipe.launch_fps(slot ^ 1, resident[(i + 1) % pool])    # device-to-device load of the next batch + its geometry
                flush.zero_()                                # evict L2 between timed iterations
                pipe.launch_main(slot)
            pipe.join()                                      # the last pyramid (and both forward streams) belong to the timed region
            b.record()
            barrier()
            t_wall1 = time.time()
            dev_ms = a.elapsed_time(b)
        last_slot = (args.steps - 1) & 1
        epe_pipe = float(pipe.out_metrics[last_slot][0].item())
        assert abs(epe_pipe - epe_last) < 1e-6, (epe_pipe, epe_last)      # same last batch, same result as the single graph
        launches = pipe.launches_per_step * args.steps

    clocks = sampler.stop(t_clock0, t_wall1) if sampler else None

    # ---------------- end to end through the public call, host buffers --------------------------
    # (the public streaming call: pinned host batches in, one EPE3D per batch out; the H2D of batch i+1 overlaps the
    # kernels of batch i, every step still moves its own inputs and reads its own result)
    e2e_runner = pipe if pipe is not None else runner
    e2e_runner.run_host_pipelined(host[i % pool] for i in range(min(2, args.warmup)))
    barrier()
    t0 = time.perf_counter()
    epes = e2e_runner.run_host_pipelined(host[i % pool] for i in range(args.steps))
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    assert len(epes) == args.steps and abs(epes[-1] - runner.run_host(host[(args.steps - 1) % pool])) < 1e-6

    # EPE3D of pair 0 of the LAST timed batch (run_host above left its flow in the runner): compared with the oracle below
    last = host[(args.steps - 1) % pool]
    flow0_pair0 = runner.out_flow[0].permute(1, 0) if runner.out_flow.shape[1] == 3 else runner.out_flow[0]
    epe_pair0 = float(torch.norm(flow0_pair0 - last["flow"][0].to(dev), dim=1).mean().item())

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks
    dev_ms, e2e_ms = t.tolist()

    h2d = sum(host[0][k].numel() * 4 for k in KEYS)
    line = None
    if rank == 0:
        hbm_peak, tc_peak, peak_src = peaks()
        ms_step = dev_ms / args.steps
        value = B * world / (ms_step * 1e-3)
        e2e_value = B * world * args.steps / (e2e_ms * 1e-3)
        kr = kernel_rooflines(torch, dev, B, hbm_peak, tc_peak)
        pc = kr["pointconv_fused"]
        cap = ncu_capture_of_roofline_kernel()
        cfg = workload_config(B)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": cfg,
            "details": {"global_pairs_per_step": B * world, "sharding": "batch-sharded, no collectives",
                        "cuda_graph": bool(graphed), "l2": "256 MB flush between timed iterations (inside the timed region "
                                                           "of the pipelined run)",
                        "pipeline": ("two streams: forward of batch i beside the sampling pyramid (4-level FPS) of batch i+1; "
                                     "K pyramids + K forwards inside the timed region; results bit-identical to the single "
                                     "graph (tests/test_runner_gpu.py)") if pipe is not None else "none (single graph per batch)",
                        "single_graph_ms_per_step": single_ms_step, "single_graph_pairs_per_s": B * world / (single_ms_step * 1e-3),
                        "pipeline_sm_limit": None if pipe is None else pipe.sm_limit,
                        "pipeline_forward_streams": None if pipe is None else (2 if pipe.dual_forward else 1),
                        "epe3d_last_step": epe_last, "epe3d_pair0_last_step": epe_pair0,
                        "parity": "FPS/kNN indices bit-exact; layer outputs 1e-4 relative on every element with shared "
                                  "kNN inputs; whole model: < 0.5 % of elements off by > 1e-4 of range (K-th-neighbour "
                                  "flips of the reference's own matmul-expansion noise), EPE3D within 1e-4 m",
                        "metrics_last_step": dict(zip(("EPE3D", "ACC3DS", "ACC3DR", "Outliers3D", "EPE2D", "ACC2D"),
                                                      [round(float(v), 6) for v in runner.out_metrics.tolist()]))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 24},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant kernel of the step (profiles/: 8 launches, largest single share): the fused PointConv
            "roofline": {"kernel": "tc_gemm_kernel<PointConvProducer<9,1>, StoreEpilogue> (fused PointConv, flow0 shape)",
                         "bound": "tensor", "achieved": pc["tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                         "frac": pc["tflops"] / tc_peak, "traffic": cap["traffic"], "peak_source": peak_src + ", burst bf16",
                         "algorithmic_flops_per_launch": pc["flops"], "algorithmic_bytes_per_launch": pc["bytes"],
                         "tensor_pipe_pct_ncu": cap["tensor_pipe_pct"],
                         "traffic_source": None if cap["source"] is None else
                         f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, parsed from {cap['source']}",
                         "note": pc["note"]},
            "kernels": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in kr.items()},
        }
        if not args.no_cpu_baseline and world == 1:
            # the oracle on pair 0 of the LAST timed batch: a bounded sample (1 warm-up + 2 timed single-pair forwards)
            # that doubles as the whole-model parity check at the benchmark shape
            pair0 = {k: v[:1].contiguous() for k, v in last.items()}
            v, sec, cores, epe_oracle = cpu_reference_forward(steps=2, warmup=1, pairs=pair0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "2 timed forwards of 1 pair (pair 0 of the last timed batch, 8192 pts) after 1 "
                                              "warm-up; oracle/layers_ref.py, torch matmul+topk kNN as the reference"}
            line["details"]["epe3d_oracle"] = epe_oracle
            line["details"]["epe3d_delta"] = abs(epe_oracle - epe_pair0)
        if not args.no_reference_gpu and world == 1:
            line["reference_gpu"] = reference_gpu_forward(torch, dev, host[0], B)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------- training workload
def run_train(args):
    """configs[3] / configs[4]: the distilTrain.py step (teacher forward under no_grad, student forward +
    backward, fused multi-scale + distillation loss, Adam), batch-sharded over ranks with ONE flat NCCL
    gradient all-reduce per step.  Reported beside the headline, not instead of it."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from kd_pointcloud_b200 import ops
    from kd_pointcloud_b200.flownet import student, teacher
    from kd_pointcloud_b200.sharding import FlatGradAllReduce
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
    from kd_pointcloud_b200.training import GraphedKDStep, kd_step

    B = args.batch
    t = teacher()
    t.load_state_dict(synthetic_state_dict(t.state_dict(), MODEL_SEED))
    s = student()
    s.load_state_dict(synthetic_state_dict(s.state_dict(), MODEL_SEED + 1))
    t, s = t.to(dev), s.to(dev)
    # N > 1: two graphs (forward+backward, optimizer) with the NCCL all-reduce issued eagerly between them
    use_graph = not args.no_graph
    from kd_pointcloud_b200.training import make_capturable_adam
    opt = make_capturable_adam(s.parameters(), lr=1e-3) if use_graph else torch.optim.Adam(s.parameters(), lr=1e-3)
    # mode='sum': the KD hint term is a SUM over the batch (loss_functions.py:213-214), so ranks add their gradients and the
    # flow terms' batch means are taken over the global batch (tests/test_multigpu_nccl.py: == single-GPU gradients)
    reducer = FlatGradAllReduce(s.parameters(), module=s, local_batch=B, mode="sum") if world > 1 else None
    kind = "kitti" if world > 1 else "ft3d"
    pool = [make_pairs(B, NPOINTS, seed=4321 + 1000 * rank + i, kind=kind, device=dev) for i in range(2)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n_step0 = ops.LAUNCHES
    kd_step(t, s, pool[0], opt, reducer)
    launches_per_step = ops.LAUNCHES - n_step0
    stepper = GraphedKDStep(t, s, pool[0], opt, reducer) if use_graph else None
    graphed = stepper is not None and stepper.graph is not None
    run = stepper.step if graphed else (lambda batch: kd_step(t, s, batch, opt, reducer))
    for i in range(args.warmup):
        run(pool[i % 2])
    barrier()
    n0 = ops.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        loss = run(pool[i % 2])
    b.record()
    barrier()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = ms.item() / args.steps
        print(json.dumps({
            "metric": "KD training pairs/sec @8192 pts", "value": B * world / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[3]/[4]: teacher fwd (no grad) + student fwd/bwd + fused KD loss + Adam; "
                                   f"{kind}-shaped synthetic 8192-pt pairs", "pairs_per_gpu_per_step": B,
                       "collective": "one flat fp32 gradient all-reduce (NCCL, eager, between the two graphs)" if world > 1 else "none",
                       "grad_elements": None if reducer is None else reducer.numel, "final_loss": float(loss.item()),
                       "cuda_graph": graphed},
            "gpu_launches": int(launches_per_step * args.steps if graphed else ops.LAUNCHES - n0)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "kd_train":
        run_train(args)
    else:
        run_kdpc(args)


if __name__ == "__main__":
    main()
