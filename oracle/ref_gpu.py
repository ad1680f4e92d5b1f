"""TEST / BASELINE INFRASTRUCTURE ONLY — never imported by kd_pointcloud_b200/.

Loads the UNMODIFIED reference scripts from baseline/_ref (tools/install_reference.sh) in one of two
ways and hands back their modules, isolated from each other and from the rest of ``sys.modules``:

  stack="compat"  models_bid_pointconv.py / models_bid_lighttoken_res.py / loss_functions.py import
                  ``pointconv_util`` / ``pointconv_util2`` / ``pointnet2.pointnet2_utils`` from compat/,
                  i.e. they run on the kdpc kernels (the drop-in route).
  stack="stock"   the same files import the reference's OWN pointconv_util.py (torch-eager layer library)
                  and its own pointnet2/pointnet2_utils.py, whose ``pointnet2_cuda`` extension is provided
                  by a ctypes module over oracle/_ref/libpointnet2_ref.so — the reference's own .cu
                  launchers compiled for sm_100a (oracle/Makefile).  This is the GPU-side baseline the
                  reference would be if somebody simply rebuilt it for Blackwell.

The ``pointnet2_cuda`` module below mirrors the pybind wrappers of pointnet2/src/pointnet2_api.cpp:10-24
(sampling.cpp:10-49, group_points.cpp, interpolate.cpp, ball_query.cpp): same names, same argument order.
"""
from __future__ import annotations

import ctypes
import os
import sys
import types
from typing import Dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpointnet2_ref.so")
COMPAT = os.path.join(ROOT, "compat")

# module names either stack binds; they are swapped in and out of sys.modules around the import
_NAMES = ("pointnet2", "pointnet2.pointnet2_utils", "pointnet2_cuda", "pointconv_util", "pointconv_util2", "pointconv_util3",
          "vn_layers", "loss_functions", "models_bid_pointconv", "models_bid_lighttoken_res", "thop", "pptk",
          "evaluation_utils")


def available() -> bool:
    return os.path.exists(os.path.join(REF_INSTALL, "models_bid_pointconv.py"))


def stock_available() -> bool:
    return available() and os.path.exists(REF_LIB)


def make_pointnet2_cuda() -> types.ModuleType:
    """ctypes twin of the reference's pybind extension, over the reference's own compiled launchers."""
    import torch
    L = ctypes.CDLL(REF_LIB)
    for n in ("ref_fps", "ref_gather", "ref_group", "ref_three_nn", "ref_three_interpolate", "ref_ball_query",
              "ref_gather_grad", "ref_group_grad", "ref_three_interpolate_grad"):
        getattr(L, n).restype = None
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    s = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    m = types.ModuleType("pointnet2_cuda")

    def furthest_point_sampling_wrapper(B, N, m_, points, temp, idx):                  # sampling.cpp:38-49
        L.ref_fps(B, N, m_, p(points), p(temp), p(idx), s())
        return 1

    def gather_points_wrapper(B, C, N, npoints, points, idx, out):                    # sampling.cpp:10-21
        L.ref_gather(B, C, N, npoints, p(points), p(idx), p(out), s())
        return 1

    def gather_points_grad_wrapper(B, C, N, npoints, grad_out, idx, grad_points):     # sampling.cpp:24-35
        L.ref_gather_grad(B, C, N, npoints, p(grad_out), p(idx), p(grad_points), s())
        return 1

    def group_points_wrapper(B, C, N, npoints, nsample, points, idx, out):            # group_points.cpp
        L.ref_group(B, C, N, npoints, nsample, p(points), p(idx), p(out), s())
        return 1

    def group_points_grad_wrapper(B, C, N, npoints, nsample, grad_out, idx, grad_points):
        L.ref_group_grad(B, C, N, npoints, nsample, p(grad_out), p(idx), p(grad_points), s())
        return 1

    def three_nn_wrapper(B, N, m_, unknown, known, dist2, idx):                        # interpolate.cpp
        L.ref_three_nn(B, N, m_, p(unknown), p(known), p(dist2), p(idx), s())

    def three_interpolate_wrapper(B, c, m_, n, points, idx, weight, out):
        L.ref_three_interpolate(B, c, m_, n, p(points), p(idx), p(weight), p(out), s())

    def three_interpolate_grad_wrapper(B, c, n, m_, grad_out, idx, weight, grad_points):
        L.ref_three_interpolate_grad(B, c, n, m_, p(grad_out), p(idx), p(weight), p(grad_points), s())

    def ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx):         # ball_query.cpp
        L.ref_ball_query(B, N, npoint, ctypes.c_float(radius), nsample, p(new_xyz), p(xyz), p(idx), s())
        return 1

    for f in (furthest_point_sampling_wrapper, gather_points_wrapper, gather_points_grad_wrapper, group_points_wrapper,
              group_points_grad_wrapper, three_nn_wrapper, three_interpolate_wrapper, three_interpolate_grad_wrapper,
              ball_query_wrapper):
        setattr(m, f.__name__, f)
    m._lib = L
    return m


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load(stack: str) -> Dict[str, types.ModuleType]:
    """Import the unmodified reference model / loss files on the given stack.  Returns a dict with the modules
    ``models_bid_pointconv``, ``models_bid_lighttoken_res``, ``loss_functions``, ``pointconv_util`` (whichever one
    the files bound).  ``sys.modules`` and ``sys.path`` are left as they were."""
    if not available():
        raise FileNotFoundError("baseline/_ref is missing: run tools/install_reference.sh in the build container")
    if stack not in ("compat", "stock"):
        raise ValueError(stack)
    saved = {k: sys.modules.pop(k) for k in _NAMES if k in sys.modules}
    saved_path = list(sys.path)
    try:
        if stack == "compat":
            sys.path[:0] = [COMPAT, ROOT, REF_INSTALL]
        else:
            if not os.path.exists(REF_LIB):
                raise FileNotFoundError("oracle/_ref/libpointnet2_ref.so is missing (built where /root/reference exists)")
            sys.modules["pointnet2_cuda"] = make_pointnet2_cuda()
            nothing = lambda *a, **k: None
            sys.modules["thop"] = _stub("thop", profile=nothing, clever_format=nothing)
            sys.modules["pptk"] = _stub("pptk")
            sys.path[:0] = [REF_INSTALL]
            import pointconv_util as R                      # the reference's own layer library
            import pointconv_util3 as R3
            R.BottleNeck = R3.BottleNeck                    # models_bid_pointconv.py:7 imports a name pointconv_util.py lacks
        import loss_functions
        import models_bid_lighttoken_res
        import models_bid_pointconv
        import pointconv_util
        out = {"models_bid_pointconv": models_bid_pointconv, "models_bid_lighttoken_res": models_bid_lighttoken_res,
               "loss_functions": loss_functions, "pointconv_util": pointconv_util,
               "pointnet2_utils": sys.modules.get("pointnet2.pointnet2_utils"),
               "pointconv_util3": sys.modules.get("pointconv_util3")}
        src = os.path.abspath(models_bid_pointconv.__file__)
        assert src.startswith(REF_INSTALL), f"models_bid_pointconv came from {src}, not from the reference install"
        return out
    finally:
        for k in _NAMES:
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        sys.path[:] = saved_path


def load_reference_pointnet2_utils(backend: str) -> types.ModuleType:
    """The reference's OWN pointnet2/pointnet2_utils.py (unmodified, from baseline/_ref) on top of a ``pointnet2_cuda``
    module: backend="stock" -> the reference's kernels (oracle/_ref), backend="kdpc" -> compat/pointnet2_cuda.py, the
    ctypes binding of libkdpc.so with the pybind wrapper names (INTEGRATION.md route B)."""
    import importlib.util
    if backend == "stock":
        cuda_mod = make_pointnet2_cuda()
    elif backend == "kdpc":
        spec = importlib.util.spec_from_file_location("pointnet2_cuda", os.path.join(COMPAT, "pointnet2_cuda.py"))
        cuda_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cuda_mod)
    else:
        raise ValueError(backend)
    saved = sys.modules.get("pointnet2_cuda")
    sys.modules["pointnet2_cuda"] = cuda_mod
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_pointnet2_utils_{backend}", os.path.join(REF_INSTALL, "pointnet2", "pointnet2_utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("pointnet2_cuda", None)
        else:
            sys.modules["pointnet2_cuda"] = saved
    return mod
