"""Per-chunk pipeline trace (clock64 stamps of CTA 0) of the fused PointConv kernel at the flow0 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import functional as KF, _lib
from kd_pointcloud_b200 import pointconv_util as P
K = torch.ops.kdpc
dev = "cuda:0"
B, N, S, k, D, Cout = 8, 8192, 8192, 9, 128, 128
torch.manual_seed(0)
cand = torch.rand(B, N, 3, device=dev) * 10
idx = K.knn(cand, cand, k)
feats = torch.randn(B, N, D, device=dev)
wn = P.WeightNet(3, 16).to(dev)
lin = torch.nn.Linear(16 * (D + 3), Cout).to(dev)
wp = K.pack_weight(lin.weight.detach(), 1, D, 16)
params = KF._weightnet_host_params(wn.mlp_convs)
L = _lib.lib()
import ctypes
L.kdpc_tc_set_trace.restype = None
L.kdpc_tc_set_trace.argtypes = [ctypes.c_void_p]
tr = torch.zeros(200 * 16, dtype=torch.int64, device=dev)
L.kdpc_tc_set_trace(tr.data_ptr())
KF._knn_compute(k, cand, cand)
order = KF.morton_order(cand) if os.environ.get("TRACE_ORDER", "1") == "1" else None
y = K.pointconv_fused(cand, cand, feats, idx, params, wp, Cout, None, lin.bias.detach(), 0.1, order)
torch.cuda.synchronize()
L.kdpc_tc_set_trace(None)
t = tr.cpu().view(200, 16)
t0 = int(t[0, 0])
print("it | w0: acq_start acquired filled arrived | w7: same | mma: wait_start full_a full_b issued  (cycles from start)")
for i in list(range(0, 40)) + list(range(60, 72)):
    r = [int(x) - t0 if int(x) else -1 for x in t[i, :16]]
    print(f"{i:3d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} | {r[4]:7d} {r[5]:7d} {r[6]:7d} {r[7]:7d} | {r[8]:7d} {r[9]:7d} {r[10]:7d} {r[11]:7d}")

import numpy as np
a = t.numpy().astype(np.int64)
rows = slice(40, 130)                      # steady state (tiles 2..4 of CTA 0)
for name, w in (("warp 0", 0), ("warp 7", 4)):
    acq = a[rows, w + 1] - a[rows, w + 0]
    fill = a[rows, w + 2] - a[rows, w + 1]
    arr = a[rows, w + 3] - a[rows, w + 2]
    per = np.diff(a[rows, w + 0])
    print(f"{name}: per chunk {per.mean():7.0f} cycles = wait for a free stage {acq.mean():6.0f} + fill {fill.mean():6.0f} + fence/arrive {arr.mean():5.0f} + rest {per.mean() - acq.mean() - fill.mean() - arr.mean():5.0f}")
mw = a[rows, 9] - a[rows, 8]
mb = a[rows, 10] - a[rows, 9]
mi = a[rows, 11] - a[rows, 10]
print(f"MMA warp: wait full_a {mw.mean():6.0f}  wait full_b {mb.mean():5.0f}  issue {mi.mean():5.0f}  per chunk {np.diff(a[rows, 8]).mean():7.0f}")
