#!/usr/bin/env python
"""Time the fused cost-volume kernel alone at the model's shapes (CUDA events, 20 launches after 3 warm-ups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
for (B, N, D) in [(8, 8192, 32), (8, 2048, 64), (8, 512, 128), (8, 256, 256)]:
    torch.manual_seed(0)
    xyz1 = torch.rand(B, N, 3, device=dev) * 10
    xyz2 = xyz1 + 0.05 * torch.randn(B, N, 3, device=dev)
    idx = K.knn(xyz1, xyz2, 32)
    p1, p2 = torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev)
    pw, pb = torch.randn(D, 3, device=dev), torch.randn(D, device=dev)
    wp = K.pack_weight(torch.randn(D, D, device=dev), 0, 0, 0)
    f = lambda: K.costvol_fused(xyz1, xyz2, p1, p2, idx, pw, pb, 0.1, wp, D, pb, 0.1)
    for _ in range(3):
        y = f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        y = f()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} N={N:5d} K=32 D={D:3d}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us   checksum {y.double().sum().item():.6e}")
