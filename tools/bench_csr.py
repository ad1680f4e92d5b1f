#!/usr/bin/env python
"""Time kdpc_build_csr at the KD step's shapes (random selections vs the real 3-NN / kNN index lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200.synth import make_pairs
K = torch.ops.kdpc
dev = "cuda:0"
def t(fn, it=20):
    fn(); fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it * 1e3
d = make_pairs(8, 8192, seed=3, device=dev)
xyz = d["pos1"]
sparse = K.gather_rows(xyz, K.fps(xyz, 2048))
cases = {"3-NN 8192<-2048 (real)": (K.knn(xyz, sparse, 3), 2048), "random (8, 8192, 3) -> 2048": (torch.randint(0, 2048, (8, 8192, 3), device=dev).int(), 2048),
         "kNN K=32 8192x8192 (real)": (K.knn(xyz, d["pos2"], 32), 8192), "random (8, 8192, 32) -> 8192": (torch.randint(0, 8192, (8, 8192, 32), device=dev).int(), 8192)}
for name, (idx, n) in cases.items():
    off, perm = K.build_csr(idx, n)
    seg = (off[:, 1:] - off[:, :-1])
    print(f"{name:32s}: {t(lambda: K.build_csr(idx, n)):8.1f} us   longest segment {int(seg.max())}, mean {float(seg.float().mean()):.1f}")
