#!/usr/bin/env python
"""One small launch of every hand-rolled-synchronisation kernel (tcgen05/TMEM/mbarrier pipelines, cluster DSMEM FPS,
warp-cooperative kNN), for `compute-sanitizer --tool memcheck|racecheck|synccheck` (tools/sanitize.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200 import pointconv_util as P
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict

dev = torch.device("cuda:0")
K = torch.ops.kdpc
which = set(sys.argv[1:]) or {"fps", "knn", "linear", "pointconv", "costvol", "misc"}
torch.manual_seed(0)
d = make_pairs(2, 4096, seed=3, device=dev)
xyz, xyz2 = d["pos1"], d["pos2"]
with torch.no_grad():
    if "fps" in which:
        a = K.fps(xyz, 512)                                    # fps_cluster_kernel (N >= 4096)
        b = K.fps(xyz[:, :1024].contiguous(), 256)             # fps_smem_kernel
        print("fps", a.shape, b.shape, int(a.sum()), int(b.sum()))
    if "knn" in which:
        for k in (3, 9, 16, 32):
            i = KF.knn_idx(k, xyz2, xyz)                       # spatial_sort_kernel + knn_bf_kernel
            print("knn", k, int(i.sum()))
        print("knn brute", int(K.knn_bruteforce(xyz[:, :300].contiguous(), xyz2, 16).sum()))
    if "linear" in which:
        x = torch.randn(2 * 4096, 128, device=dev)
        for n, kk in ((128, 128), (64, 128), (256, 128), (32, 64)):
            w = torch.randn(n, kk, device=dev)
            y = KF.fused_linear(x[:, :kk].contiguous(), w, torch.randn(n, device=dev), None, 0.1)    # PlainAsyncProducer / PlainProducer
            print("linear", n, kk, float(y.abs().sum()))
        y = KF.fused_linear(torch.randn(256, 2096, device=dev), torch.randn(128, 2096, device=dev))   # split-K + splitk_reduce
        print("linear split-K", float(y.abs().sum()))
        y = KF.fused_linear(x[:, :3].contiguous(), torch.randn(32, 3, device=dev))                    # linear_simt
        print("linear simt", float(y.abs().sum()))
    if "pointconv" in which:
        for ksz, cin, cout, npoint in ((9, 64, 128, None), (16, 64, 64, 512)):
            layer = P.PointConv(ksz, cin + 3, cout, bn=True) if npoint is None else P.PointConvD(npoint, ksz, cin + 3, cout)
            layer.load_state_dict(synthetic_state_dict(layer.state_dict(), 1))
            layer = layer.to(dev).eval()
            out = layer(xyz.permute(0, 2, 1), torch.randn(2, cin, 4096, device=dev))
            out = out if torch.is_tensor(out) else out[1]
            print("pointconv", ksz, float(out.abs().sum()))    # PointConvProducer<9,1> / <16,2> (+ weightnet pre-pass)
    if "costvol" in which:
        for dch in (32, 256):
            cl = P.CrossLayerLight(32, 48, [dch, dch], [dch, dch])
            cl.load_state_dict(synthetic_state_dict(cl.state_dict(), 2))
            cl = cl.to(dev).eval()
            n = 4096 if dch == 32 else 512
            o = cl(xyz[:, :n].permute(0, 2, 1), xyz2[:, :n].permute(0, 2, 1), torch.randn(2, 48, n, device=dev), torch.randn(2, 48, n, device=dev))
            print("costvol", dch, float(o[2].abs().sum()))     # CostVolAsyncProducer / CostVolProducer + MaxKEpilogue
    if "misc" in which:
        idx = KF.knn_idx(3, xyz[:, :1024].contiguous(), xyz)
        print("interp3", float(KF.interp3(xyz, xyz[:, :1024].contiguous(), idx, torch.randn(2, 1024, 64, device=dev)).abs().sum()))
torch.cuda.synchronize()
print("sanitize_small done")
