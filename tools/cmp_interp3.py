#!/usr/bin/env python
"""interp3 at the model's shapes: sha256 of the outputs (compare two builds with KDPC_LIB=...) and L2-flushed CUDA-event times."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kd_pointcloud_b200 import ops  # noqa: F401
K = torch.ops.kdpc
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (B, N, S, C) in [(8, 8192, 2048, 64), (16, 8192, 2048, 32), (8, 8192, 2048, 3), (8, 2048, 512, 128), (16, 512, 256, 256), (3, 1000, 77, 20), (2, 333, 50, 5), (1, 1, 3, 4)]:
    g = torch.Generator().manual_seed(B * N + C)
    q = (torch.rand(B, N, 3, generator=g) * 10).to(dev)
    c = (torch.rand(B, S, 3, generator=g) * 10).to(dev)
    f = torch.randn(B, S, C, generator=g).to(dev)
    idx = K.knn(q, c, 3)
    out, w = K.interp3(q, c, idx, f)
    h = hashlib.sha256(out.cpu().numpy().tobytes() + w.cpu().numpy().tobytes()).hexdigest()[:16]
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.interp3(q, c, idx, f); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"B={B:2d} N={N:5d} S={S:5d} C={C:3d}: {ts[3]:7.1f} us  sha {h}")
