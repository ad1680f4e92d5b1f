#!/usr/bin/env python
"""Launch the training-path kernels once at the student's shapes (for `ncu -k regex:<name>` captures):
agg_grad (warp-per-point aggregation backward), costvol_grad (arg-max cost-volume backward), dw (tcgen05 / CUDA-core dW)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
what = sys.argv[1:] or ["agg_grad", "costvol_grad", "dw"]
torch.manual_seed(0)
for _ in range(2):
    if "agg_grad" in what:
        rows, k, c = 65536, 9, 131
        K.pointconv_agg_grad(torch.randn(1, rows, k, c, device=dev), torch.rand(1, rows, k, 16, device=dev),
                             torch.randn(1, rows, c * 16, device=dev), True, True)
    if "costvol_grad" in what:
        B, N, D = 8, 8192, 32
        xyz = torch.rand(B, N, 3, device=dev) * 10
        idx = K.knn(xyz, xyz + 0.05 * torch.randn_like(xyz), 32)
        K.costvol_grad(torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev), idx, torch.randn(D, D, device=dev) / 6,
                       torch.randn(D, device=dev), 0.1, 0.1, torch.randn(B, N, D, device=dev))
    if "dw" in what:
        for m, n, k in ((2097152, 32, 32), (2097152, 32, 3), (65536, 128, 2096)):
            K.linear_dw(torch.randn(m, n, device=dev), torch.randn(m, k, device=dev), True)
torch.cuda.synchronize()
print("ok")
