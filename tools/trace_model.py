#!/usr/bin/env python
"""Per-call device time of one eager forward (B=8, 8192 points): which C-ABI call, which shape, how long."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import ops, functional as KF
from kd_pointcloud_b200.flownet import PointConvBidirection
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
dev = "cuda:0"
m = PointConvBidirection()
m.load_state_dict(synthetic_state_dict(m.state_dict(), 7))
m = m.to(dev).eval()
d = make_pairs(8, 8192, seed=1234, device=dev)
with torch.no_grad():
    for _ in range(2):
        KF.clear_caches(); m(d["pos1"], d["pos2"], d["color1"], d["color2"])
    torch.cuda.synchronize()
    KF.clear_caches()
    ops.TRACE = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); m(d["pos1"], d["pos2"], d["color1"], d["color2"]); t1.record()
    torch.cuda.synchronize()
tr, ops.TRACE = ops.TRACE, None
rows = [(n, a, s.elapsed_time(e) * 1e3) for n, a, s, e in tr]
print(f"eager forward {t0.elapsed_time(t1):.2f} ms, {len(rows)} kdpc calls, sum of call times {sum(r[2] for r in rows)/1e3:.2f} ms")
by = collections.defaultdict(lambda: [0, 0.0])
for n, a, t in rows:
    by[n][0] += 1; by[n][1] += t
for n, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us  n={c:3d}  {n}")
print("--- calls > 40 us, in order")
for n, a, t in rows:
    if t > 40: print(f"{t:8.1f} us  {n:24s} {a}")
print("--- linear calls (m, n, k, ...), in order")
for n, a, t in rows:
    if n in ("kdpc_linear_tc", "kdpc_linear_simt"): print(f"{t:8.1f} us  {n:18s} {a[:4]}")
