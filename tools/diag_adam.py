#!/usr/bin/env python
"""Which parameters does KdpcAdam update differently from torch.optim.Adam after ONE KD backward? (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import flownet, training
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
dev = "cuda:0"
teacher = flownet.teacher().to(dev)
teacher.load_state_dict(synthetic_state_dict(teacher.state_dict(), 0))
batch = {k: v.to(dev) for k, v in make_pairs(4, 4096, seed=3).items()}
out = {}
for name in ("plain", "kdpc"):
    student = flownet.student().to(dev)
    student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
    ps = list(student.parameters())
    opt = torch.optim.Adam(ps, lr=1e-3) if name == "plain" else training.KdpcAdam(ps, lr=1e-3)
    w0 = [p.detach().clone() for p in ps]
    training.kd_step(teacher, student, batch, opt)
    out[name] = (w0, [p.detach().clone() for p in ps], [None if p.grad is None else p.grad.detach().clone() for p in ps],
                 [n for n, _ in student.named_parameters()], [None if p.grad is None else (p.grad.data_ptr(), p.grad.is_contiguous(), tuple(p.grad.stride())) for p in ps])
w0, wa, ga, names, meta = out["plain"]
_, wb, gb, _, metab = out["kdpc"]
ptrs = {}
for n, m in zip(names, metab):
    if m is not None:
        ptrs.setdefault(m[0], []).append(n)
print("aliased grads:", {k: v for k, v in ptrs.items() if len(v) > 1})
bad = 0
for n, a, b, x, y, m in zip(names, wa, wb, ga, gb, metab):
    if x is None:
        continue
    dg = (x - y).abs().max().item()
    dw = (a - b).abs().max().item()
    if dw > 1e-6 or dg > 0:
        bad += 1
        if bad < 15:
            print(f"{n:50s} |dW| {dw:.3e}  |dgrad| {dg:.3e}  step {(a - w0[names.index(n)]).abs().max().item():.3e} vs {(b - w0[names.index(n)]).abs().max().item():.3e}  {m[1:]}")
print("parameters with different updates:", bad, "of", sum(x is not None for x in ga))
