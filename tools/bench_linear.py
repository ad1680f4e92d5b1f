import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
K = torch.ops.kdpc
dev = "cuda:0"
def t(fn, it=50):
    """device time per call: `it` calls captured in one CUDA graph (a Python call costs ~20 us of host time, more than
    the small layers take on the device)"""
    fn(); fn(); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(it): fn()
    g_.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g_.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / it * 1e3
for (m, n, k) in [(128, 32, 32), (128, 256, 256), (2048, 256, 256), (4096, 256, 256), (4096, 256, 320), (8192, 256, 128), (8192, 128, 192), (4096, 128, 128),
                  (2048, 128, 128), (4096, 64, 256), (2048, 64, 64), (16384, 64, 64), (131072, 32, 32), (131072, 64, 32),
                  (65536, 128, 128), (65536, 64, 128), (131072, 64, 64), (262144, 64, 64)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev)
    wp = K.pack_weight(w, 0, 0, 0)
    sh = torch.randn(n, device=dev)
    us = t(lambda: K.linear_tc(x, wp, n, None, sh, 0.1, 1.0, 0.0, None))
    us_t = t(lambda: torch.nn.functional.leaky_relu(torch.nn.functional.linear(x, w, sh), 0.1))
    byts = (m * k + m * n) * 4
    print(f"M={m:7d} N={n:3d} K={k:4d}: linear_tc {us:7.1f} us ({byts/us/1e3:7.1f} GB/s)   torch sgemm+act {us_t:7.1f} us")
