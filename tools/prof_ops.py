#!/usr/bin/env python
"""Launch individual kdpc kernels at model shapes (for `ncu -k regex:<name>` captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200.synth import make_pairs
K = torch.ops.kdpc
what = sys.argv[1:] or ["knn"]
dev = "cuda:0"
d = make_pairs(8, 8192, seed=99, device=dev)
xyz, xyz2 = d["pos1"], d["pos2"]
for _ in range(2):
    if "knn" in what:
        for k in (32, 9, 3):
            K.knn_bruteforce(xyz2, xyz, k)
            qs, cs = K.spatial_sort(xyz2), K.spatial_sort(xyz)
            K.knn_sorted(qs, cs, 8, 8192, 8192, k)
    if "costvol" in what:
        N, D = 8192, 32
        idx32 = K.knn(xyz, xyz2, 32)
        p1, p2 = torch.randn(8, N, D, device=dev), torch.randn(8, N, D, device=dev)
        pw, pb = torch.randn(D, 3, device=dev), torch.randn(D, device=dev)
        wp = K.pack_weight(torch.randn(D, D, device=dev), 0, 0, 0)
        from kd_pointcloud_b200 import _lib
        for mode in (1, 0):
            _lib.lib().kdpc_tc_set_async(mode)
            K.costvol_fused(xyz, xyz2, p1, p2, idx32, pw, pb, 0.1, wp, D, pb, 0.1)
        _lib.lib().kdpc_tc_set_async(1)
    if "pointconv" in what:
        from kd_pointcloud_b200 import pointconv_util as P
        D, Cout = 128, 128
        idx9 = KF._knn_compute(9, xyz, xyz)            # (leaves the Morton order in the sort cache)
        feats = torch.randn(8, 8192, D, device=dev)
        wn = P.WeightNet(3, 16).to(dev)
        lin = torch.nn.Linear(16 * (D + 3), Cout).to(dev)
        wp = K.pack_weight(lin.weight.detach(), 1, D, 16)
        params = KF._weightnet_host_params(wn.mlp_convs)
        K.pointconv_fused(xyz, xyz, feats, idx9, params, wp, Cout, None, lin.bias.detach(), 0.1, KF.morton_order(xyz))
    if "linear" in what:
        x = torch.randn(65536, 2096, device=dev)
        w = torch.randn(128, 2096, device=dev)
        wp = K.pack_weight(w, 0, 0, 0)
        K.linear_tc(x, wp, 128, None, None, 0.1, 1.0, 0.0, None)
        x3 = torch.randn(128, 256, device=dev)
        wp3 = K.pack_weight(torch.randn(256, 256, device=dev), 0, 0, 0)
        K.linear_tc(x3, wp3, 256, None, None, 0.1, 1.0, 0.0, None)
        x2 = torch.randn(8 * 8192 * 4, 64, device=dev)
        wp2 = K.pack_weight(torch.randn(64, 64, device=dev), 0, 0, 0)
        K.linear_tc(x2, wp2, 64, None, None, 0.1, 1.0, 0.0, None)
    if "fps" in what:
        K.fps(xyz, 2048)
torch.cuda.synchronize()
print("ok")
