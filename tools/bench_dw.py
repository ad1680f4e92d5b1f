#!/usr/bin/env python
"""Time the tcgen05 weight gradient (kdpc_linear_dw) at the student's training shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
for m, n, k in [(2097152, 32, 32), (2097152, 32, 3), (524288, 64, 64), (524288, 64, 3), (131072, 128, 128), (65536, 256, 256), (131072, 32, 64),
                (131072, 128, 3), (65536, 128, 2096), (16384, 128, 3120)]:
    dy, x = torch.randn(m, n, device=dev), torch.randn(m, k, device=dev)
    for _ in range(3): K.linear_dw(dy, x, True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): K.linear_dw(dy, x, True)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 * 1e3
    print(f"m={m:8d} n={n:3d} k={k:4d}: {us:8.1f} us  {4 * m * (n + k) / us / 1e3:7.1f} GB/s")
