#!/usr/bin/env python
"""torch's fused vs foreach capturable Adam on the student's real gradients: parameter deltas after 1 and 3 eager steps."""
import os, sys, copy
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
res = {}
for fused in (False, True):
    student = flownet.student().to(dev)
    student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
    opt = training.make_capturable_adam(student.parameters(), lr=1e-3, fused=fused)
    losses = [float(training.kd_step(teacher, student, batch, opt)) for _ in range(3)]
    res[fused] = (losses, [p.detach().clone() for p in student.parameters()])
print("losses foreach", res[False][0], "fused", res[True][0])
worst = max(((a - b).abs().max() / (a.abs().max() + 1e-12)).item() for a, b in zip(res[False][1], res[True][1]))
print("max relative parameter difference after 3 steps:", worst)
