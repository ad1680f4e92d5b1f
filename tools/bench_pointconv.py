#!/usr/bin/env python
"""Time the fused PointConv kernel alone at the model's shapes (CUDA events, 20 launches after 3 warm-ups)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import functional as KF, _lib
from kd_pointcloud_b200 import pointconv_util as P
K = torch.ops.kdpc
dev = "cuda:0"
# (B, N, S, K, D, Cout): flow0..flow3 PointConv #1/#2, encoder PointConvD l1..l4 (2B clouds)
shapes = [(8, 8192, 8192, 9, 128, 128), (8, 2048, 2048, 9, 192, 128), (8, 2048, 2048, 9, 128, 128), (8, 512, 512, 9, 320, 128),
          (16, 8192, 2048, 16, 64, 64), (16, 2048, 512, 16, 128, 128), (16, 512, 256, 16, 256, 256), (16, 256, 64, 16, 512, 256)]
for (B, N, S, k, D, Cout) in shapes:
    torch.manual_seed(0)
    cand = torch.rand(B, N, 3, device=dev) * 10
    query = cand[:, torch.randperm(N, device=dev)[:S]].contiguous() if S != N else cand
    idx = K.knn(query, cand, k)
    feats = torch.randn(B, N, D, device=dev)
    wn = P.WeightNet(3, 16).to(dev)
    lin = torch.nn.Linear(16 * (D + 3), Cout).to(dev)
    wp = K.pack_weight(lin.weight.detach(), 1, D, 16)
    params = KF._weightnet_host_params(wn.mlp_convs)
    out = {}
    KF._knn_compute(k, cand, query)                       # leaves the Morton order of the queries in the sort cache
    mo = KF.morton_order(query)
    for staged in (0, 1, 2):
        order = mo if staged >= 1 else None
        _lib.lib().kdpc_pointconv_set_precompute(0 if staged == 2 else 1)
        for _ in range(3):
            y = K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, None, lin.bias.detach(), 0.1, order)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            y = K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, None, lin.bias.detach(), 0.1, order)
        e1.record()
        torch.cuda.synchronize()
        out[staged] = (e0.elapsed_time(e1) / 20 * 1e3, y)
    same = torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][1], out[2][1])
    fl = 2.0 * B * S * (D + 3) * 16 * (k + Cout)
    print(f"B={B:2d} N={N:5d} S={S:5d} K={k:2d} D={D:3d} Cout={Cout:3d}: natural order {out[0][0]:7.1f} us   "
          f"Morton order {out[1][0]:7.1f} us ({fl / out[1][0] / 1e6:6.1f} TF/s)   Morton, WeightNet in-kernel {out[2][0]:7.1f} us   identical={same}")
_lib.lib().kdpc_pointconv_set_precompute(1)
