#!/usr/bin/env python
"""Time the PointConv aggregation backward (kdpc_pointconv_agg_grad) at the student's training shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
for rows, k, c in [(65536, 9, 131), (32768, 16, 67), (16384, 9, 195), (16384, 9, 131), (8192, 16, 131), (4096, 16, 259), (4096, 9, 323), (2048, 9, 515), (1024, 16, 515)]:
    grouped = torch.randn(1, rows, k, c, device=dev)
    wn = torch.rand(1, rows, k, 16, device=dev)
    go = torch.randn(1, rows, c * 16, device=dev)
    for _ in range(3): out = K.pointconv_agg_grad(grouped, wn, go, True, True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): out = K.pointconv_agg_grad(grouped, wn, go, True, True)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 * 1e3
    byts = 4 * rows * (2 * k * c + 2 * k * 16 + c * 16)
    print(f"rows={rows:6d} k={k:2d} c={c:3d}: {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s  {4 * rows * k * c * 16 / us / 1e6:6.2f} TFLOP/s")
