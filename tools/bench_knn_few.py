#!/usr/bin/env python
"""K = 3 search: one warp per query vs one thread per query with a shared tile walk per warp (kdpc_knn_set_few)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import _lib
from kd_pointcloud_b200.synth import make_pairs
K = torch.ops.kdpc
dev = "cuda:0"
d = make_pairs(8, 8192, seed=99, device=dev)
q, c = d["pos2"], d["pos1"]
qs, cs = K.spatial_sort(q), K.spatial_sort(c)
for few in (0, 1):
    _lib.lib().kdpc_knn_set_few(few)
    for k in (3, 1):
        for _ in range(3): r = K.knn_sorted(qs, cs, 8, 8192, 8192, k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(10): r = K.knn_sorted(qs, cs, 8, 8192, 8192, k)
        b.record(); torch.cuda.synchronize()
        print(f"few={few} K={k}: {a.elapsed_time(b) / 10 * 1e3:8.1f} us")
_lib.lib().kdpc_knn_set_few(0)
