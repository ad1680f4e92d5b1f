import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import _lib
from kd_pointcloud_b200.synth import make_pairs
K = torch.ops.kdpc
L = _lib.lib()
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it * 1e3
for B in (1, 8, 16, 17, 32):
    for n, m in ((8192, 2048), (2048, 512)):
        xyz = make_pairs(B, n, seed=3, device="cuda:0")["pos1"]
        L.kdpc_fps_set_cluster(1); tc = t(lambda: K.fps(xyz, m))
        L.kdpc_fps_set_cluster(0); ts = t(lambda: K.fps(xyz, m))
        L.kdpc_fps_set_cluster(1)
        print(f"B={B:3d} n={n} m={m}: cluster {tc:8.1f} us   single {ts:8.1f} us")
