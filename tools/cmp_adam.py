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
makers = {
    "torch.optim.Adam (plain: what distilTrain.py:134 builds)": lambda ps: torch.optim.Adam(ps, lr=1e-3),
    "torch Adam capturable, tensor lr (default impl)": lambda ps: torch.optim.Adam(ps, lr=torch.tensor(1e-3, device=dev), capturable=True),
    "torch Adam capturable, tensor lr, fused=True": lambda ps: torch.optim.Adam(ps, lr=torch.tensor(1e-3, device=dev), capturable=True, fused=True),
    "KdpcAdam (csrc/adam.cu)": lambda ps: training.KdpcAdam(ps, lr=1e-3),
}
for name, mk in makers.items():
    student = flownet.student().to(dev)
    student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
    opt = mk(list(student.parameters()))
    losses = [round(float(training.kd_step(teacher, student, batch, opt)), 3) for _ in range(3)]
    res[name] = (losses, [p.detach().clone() for p in student.parameters()])
    print(f"{name:60s} losses {losses}")
base = res["torch.optim.Adam (plain: what distilTrain.py:134 builds)"][1]
for name, (_, ps) in res.items():
    worst = max(((a - b).abs().max() / (a.abs().max() + 1e-12)).item() for a, b in zip(base, ps))
    print(f"{name:60s} max relative parameter difference to plain Adam after 3 steps: {worst:.3e}")
