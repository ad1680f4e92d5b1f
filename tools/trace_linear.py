"""Per-iteration pipeline trace (clock64 stamps of CTA 0) of the streaming tcgen05 linear layer.
usage: trace_linear.py [M K N]   (default 262144 64 64)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import _lib
K = torch.ops.kdpc
dev = "cuda:0"
M, Kd, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (262144, 64, 64)
torch.manual_seed(0)
x = torch.randn(M, Kd, device=dev)
wp = K.pack_weight(torch.randn(N, Kd, device=dev), 0, 0, 0)
sh = torch.randn(N, device=dev)
f = lambda: K.linear_tc(x, wp, N, None, sh, 0.1, 1.0, 0.0, None)
f(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): f()
b.record(); b.synchronize()
print(f"M={M} K={Kd} N={N}: {a.elapsed_time(b) * 100:.1f} us per launch (warm L2)")
L = _lib.lib()
L.kdpc_tc_set_trace.restype = None
L.kdpc_tc_set_trace.argtypes = [ctypes.c_void_p]
tr = torch.zeros(200 * 16, dtype=torch.int64, device=dev)
L.kdpc_tc_set_trace(tr.data_ptr())
f(); torch.cuda.synchronize()
L.kdpc_tc_set_trace(None)
t = tr.cpu().view(200, 16)
t0 = int(t[0, 0])
print("it | producer t0: top synced issued stage_free raw_ready converted arrived | mma: tile_top tmem_free wait_start full_a full_b issued | epilogue q1: wait_start tmem_full done")
n_it = int((t[:, 0] != 0).sum())
for i in list(range(0, min(12, n_it))) + list(range(max(12, n_it - 6), n_it)):
    r = [int(v) - t0 if int(v) else -1 for v in t[i, :16]]
    print(f"{i:3d} | " + " ".join(f"{v:7d}" for v in r[0:7]) + " | " + f"{r[7]:7d} {r[15]:7d} " + " ".join(f"{v:7d}" for v in r[8:12]) + " | " + " ".join(f"{v:7d}" for v in r[12:15]))
if n_it > 6:
    per = (int(t[n_it - 1, 0]) - int(t[2, 0])) / (n_it - 3)
    print(f"iterations traced {n_it}; steady state {per:.0f} cycles per pipeline iteration")
