#!/usr/bin/env python
"""Time the recomputing arg-max backward of the fused cost volume (kdpc_costvol_grad) at the student's level-0 / level-1 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
for B, N, D in ((8, 8192, 32), (8, 2048, 64)):
    xyz = torch.rand(B, N, 3, device=dev) * 10
    idx = K.knn(xyz, xyz + 0.05 * torch.randn_like(xyz), 32)
    args = (torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev), idx, torch.randn(D, D, device=dev) / D ** 0.5,
            torch.randn(D, device=dev), 0.1, 0.1, torch.randn(B, N, D, device=dev))
    for _ in range(3): K.costvol_grad(*args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): K.costvol_grad(*args)
    b.record(); torch.cuda.synchronize()
    print(f"B={B} N={N} D={D}: {a.elapsed_time(b) / 10 * 1e3:8.1f} us")
