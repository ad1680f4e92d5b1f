"""Route B of INTEGRATION.md as a real module: a drop-in for the reference's pybind extension ``pointnet2_cuda``
(pointnet2/src/pointnet2_api.cpp:10-24) that binds libkdpc.so through ctypes with the SAME wrapper names and argument
order (sampling.cpp:10-49, group_points.cpp, interpolate.cpp, ball_query.cpp).  With this file importable as
``pointnet2_cuda`` the reference's own ``pointnet2/pointnet2_utils.py`` runs unmodified on the kdpc kernels.

The grad entry points OVERWRITE ``grad_points`` (the reference pre-zeroes and accumulates atomically; pre-zeroing is
harmless), results are deterministic, and launch failures raise instead of ``exit(-1)``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("KDPC_LIB") or os.path.join(_HERE, "..", "kd_pointcloud_b200", "libkdpc.so")
if not os.path.exists(_LIB):
    raise ImportError(f"{_LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
_L = ctypes.CDLL(_LIB)
_L.kdpc_error_string.restype = ctypes.c_char_p
_L.kdpc_error_string.argtypes = [ctypes.c_int]
_L.kdpc_knn_workspace_bytes.restype = ctypes.c_longlong
_L.kdpc_knn_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ok(rc, what):
    if rc:
        raise RuntimeError(f"{what}: {_L.kdpc_error_string(rc).decode()}")


def _ws(n_int32, like):            # the ABI is caller-allocates, like the reference (pointnet2_utils.py:25-26)
    return torch.empty(n_int32, dtype=torch.int32, device=like.device)


def furthest_point_sampling_wrapper(B, N, m, points, temp, idx):                 # sampling.cpp:38
    _ok(_L.kdpc_fps(B, N, m, _p(points), _p(temp), _p(idx), _s()), "fps")
    return 1


def gather_points_wrapper(B, C, N, npoints, points, idx, out):                   # sampling.cpp:10
    _ok(_L.kdpc_gather(B, C, N, npoints, _p(points), _p(idx), _p(out), _s()), "gather")
    return 1


def gather_points_grad_wrapper(B, C, N, npoints, grad_out, idx, grad_points):    # sampling.cpp:24
    ws = _ws(B * (N + 1 + npoints), grad_out)
    _ok(_L.kdpc_gather_grad(B, C, N, npoints, _p(grad_out), _p(idx), _p(ws), _p(grad_points), _s()), "gather_grad")
    return 1


def group_points_wrapper(B, C, N, npoints, nsample, points, idx, out):           # group_points.cpp
    _ok(_L.kdpc_group(B, C, N, npoints, nsample, _p(points), _p(idx), _p(out), _s()), "group")
    return 1


def group_points_grad_wrapper(B, C, N, npoints, nsample, grad_out, idx, grad_points):
    ws = _ws(B * (N + 1 + npoints * nsample), grad_out)
    _ok(_L.kdpc_group_grad(B, C, N, npoints, nsample, _p(grad_out), _p(idx), _p(ws), _p(grad_points), _s()), "group_grad")
    return 1


def three_nn_wrapper(B, N, m, unknown, known, dist2, idx):                       # interpolate.cpp
    ws = torch.empty(_L.kdpc_knn_workspace_bytes(B, N, m), dtype=torch.uint8, device=known.device)
    _ok(_L.kdpc_three_nn(B, N, m, _p(unknown), _p(known), _p(ws), _p(dist2), _p(idx), _s()), "three_nn")


def three_interpolate_wrapper(B, c, m, n, points, idx, weight, out):
    _ok(_L.kdpc_three_interpolate(B, c, m, n, _p(points), _p(idx), _p(weight), _p(out), _s()), "three_interpolate")


def three_interpolate_grad_wrapper(B, c, n, m, grad_out, idx, weight, grad_points):
    ws = _ws(B * (m + 1 + 3 * n), grad_out)
    _ok(_L.kdpc_three_interpolate_grad(B, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(ws), _p(grad_points), _s()), "three_interpolate_grad")


def ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx):        # ball_query.cpp
    _ok(_L.kdpc_ball_query(B, N, npoint, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s()), "ball_query")
    return 1
