#!/usr/bin/env python
"""One KD training step under the torch profiler: per-kernel CUDA time table (top 40)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import flownet, training
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    training.kd_step(teacher, student, batch, opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
