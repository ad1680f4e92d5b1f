#!/usr/bin/env python
"""Per-call device time of one eager KD training step (B=8, 8192 points): which C-ABI call, which shape, how long."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import flownet, training, ops
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
dev = "cuda:0"
torch.manual_seed(0)
teacher, student = flownet.teacher().to(dev), flownet.student().to(dev)
teacher.load_state_dict(synthetic_state_dict(teacher.state_dict(), 0))
student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
opt = torch.optim.Adam(student.parameters(), lr=1e-4)
batch = {k: v.to(dev) for k, v in make_pairs(8, 8192, seed=3).items()}
for _ in range(2):
    training.kd_step(teacher, student, batch, opt)
torch.cuda.synchronize()
ops.TRACE = []
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); training.kd_step(teacher, student, batch, opt); t1.record()
torch.cuda.synchronize()
tr, ops.TRACE = ops.TRACE, None
rows = [(n, a, s.elapsed_time(e) * 1e3) for n, a, s, e in tr]
print(f"eager KD step {t0.elapsed_time(t1):.2f} ms, {len(rows)} kdpc calls, sum of call times {sum(r[2] for r in rows)/1e3:.2f} ms")
by = collections.defaultdict(lambda: [0, 0.0])
for n, a, t in rows:
    by[n][0] += 1; by[n][1] += t
for n, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us  n={c:3d}  {n}")
print("--- per (call, shape), summed, > 150 us")
bs = collections.defaultdict(lambda: [0, 0.0])
for n, a, t in rows:
    bs[(n, a[:6])][0] += 1; bs[(n, a[:6])][1] += t
for (n, a), (c, t) in sorted(bs.items(), key=lambda kv: -kv[1][1]):
    if t > 150: print(f"{t:9.1f} us  n={c:3d}  {n:26s} {a}")
