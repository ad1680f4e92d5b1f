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
    if "fps" in what:
        K.fps(xyz, 2048)
torch.cuda.synchronize()
print("ok")
