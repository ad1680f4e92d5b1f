"""Per-iteration pipeline trace (clock64 stamps of CTA 0) of the fused cost-volume kernel at the cross0 shape."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from kd_pointcloud_b200 import _lib
K = torch.ops.kdpc
dev = "cuda:0"
B, N, D = 8, 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 32
if D != 32:
    N = {64: 2048, 128: 512, 256: 256}[D]
torch.manual_seed(0)
xyz1 = torch.rand(B, N, 3, device=dev) * 10
xyz2 = xyz1 + 0.05 * torch.randn(B, N, 3, device=dev)
idx = K.knn(xyz1, xyz2, 32)
p1, p2 = torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev)
pw, pb = torch.randn(D, 3, device=dev), torch.randn(D, device=dev)
wp = K.pack_weight(torch.randn(D, D, device=dev), 0, 0, 0)
f = lambda: K.costvol_fused(xyz1, xyz2, p1, p2, idx, pw, pb, 0.1, wp, D, pb, 0.1)
f(); torch.cuda.synchronize()
L = _lib.lib()
L.kdpc_tc_set_trace.restype = None
L.kdpc_tc_set_trace.argtypes = [ctypes.c_void_p]
tr = torch.zeros(200 * 16, dtype=torch.int64, device=dev)
L.kdpc_tc_set_trace(tr.data_ptr())
f(); torch.cuda.synchronize()
L.kdpc_tc_set_trace(None)
t = tr.cpu().view(200, 16)
t0 = int(t[0, 0])
print("it | producer w0: top synced issued stage_free raw_ready converted arrived | mma: tile_top tmem_free wait_start full_a full_b issued | epilogue q1: wait_start tmem_full done")
for i in list(range(0, 8)) + list(range(40, 60)):
    r = [int(x) - t0 if int(x) else -1 for x in t[i, :16]]
    print(f"{i:3d} | " + " ".join(f"{v:7d}" for v in r[0:7]) + " | " + f"{r[7]:7d} {r[15]:7d} " + " ".join(f"{v:7d}" for v in r[8:12]) + " | " + " ".join(f"{v:7d}" for v in r[12:15]))
