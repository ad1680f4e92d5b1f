#!/usr/bin/env python
"""One launch of the tcgen05 weight-gradient kernel at a chosen shape (for `ncu -k regex:dw_tc_kernel`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
m, n, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 128, 2096)))
dy, x = torch.randn(m, n, device="cuda:0"), torch.randn(m, k, device="cuda:0")
for _ in range(3): K.linear_dw(dy, x, True)
torch.cuda.synchronize()
print("ok")
