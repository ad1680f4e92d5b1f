"""FPS variants on the model's shapes: automatic choice, one-CTA kernel, and forced cluster shapes
(ctas x threads per cloud, optional spread placement).  Every variant must return the same indices."""
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
for ctas, tot in ((8, 2048), (4, 2048), (8, 1024), (4, 1024), (2, 1024)):
    print(f"capacity ctas={ctas} threads={tot} n=8192: {L.kdpc_fps_cluster_capacity(ctas, tot, 8192)} clusters")
for B in (8, 16):
    for n, m in ((8192, 2048), (2048, 512)):
        xyz = make_pairs(B, n, seed=3, device="cuda:0")["pos1"]
        L.kdpc_fps_set_cluster(0); ref = K.fps(xyz, m); ts = t(lambda: K.fps(xyz, m))
        L.kdpc_fps_set_cluster(1); ta = t(lambda: K.fps(xyz, m))
        line = f"B={B:3d} n={n} m={m}: single {ts:8.1f} us  auto {ta:8.1f} us ({ta / (m - 1):.3f} us/iter)"
        for spread in (0, 1):
            for ctas, tot in ((8, 2048), (4, 2048), (8, 1024), (4, 1024), (2, 1024), (2, 2048)):
                if B * ctas > 148 or tot > n:
                    continue
                if L.kdpc_fps_cluster_capacity(ctas, tot, n) < B:
                    line += f"  [{'s' if spread else ''}{ctas}x{tot // ctas}: no fit]"
                    continue
                L.kdpc_fps_set_cluster(spread * 1000000 + ctas * 10000 + tot)
                try:
                    out = K.fps(xyz, m)
                    ok = torch.equal(out, ref)
                    tv = t(lambda: K.fps(xyz, m))
                    line += f"  [{'s' if spread else ''}{ctas}x{tot // ctas}: {tv:7.1f} us{'' if ok else ' MISMATCH'}]"
                except Exception as e:
                    line += f"  [{ctas}x{tot // ctas}: {type(e).__name__}]"
                    torch.cuda.synchronize()
        L.kdpc_fps_set_cluster(1)
        print(line, flush=True)
