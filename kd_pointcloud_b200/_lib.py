"""ctypes binding of libkdpc.so (the C ABI declared in include/kdpc.h) and its build recipe.

The library is built IN-TREE next to this file (``build()``), so the ``.so`` travels with the
repository snapshot.  There is no CPU implementation behind this module: if the library is
missing, or a tensor is not on a CUDA device, the ops raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("KDPC_LIB") or os.path.join(_HERE, "libkdpc.so")   # KDPC_LIB: an experiment build (A/B measurements)
SOURCES = ["abi.cu", "fps.cu", "knn.cu", "group.cu", "interp.cu", "pointconv.cu", "costvol.cu",
           "scatter.cu", "metrics.cu", "linear_tc.cu", "pointconv_tc.cu", "costvol_tc.cu", "knn_bf.cu", "loss.cu", "knn_feat.cu", "dataprep.cu", "dw_tc.cu", "weightnet_grad.cu", "costvol_grad.cu", "adam.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
LINK_FLAGS = ARCH_FLAGS + ["-shared"]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(_HERE, "..", "include", "kdpc.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every kernel for sm_100a with nvcc (cross-compiles without a GPU): one object per
    source (in parallel, rebuilt only when stale) under csrc/build/, then one link."""
    force = force or os.environ.get("KDPC_FORCE_BUILD", "0") == "1"      # rebuild every object from clean
    if os.environ.get("KDPC_LIB") or (not force and not needs_build()):
        return LIB_PATH
    log = []
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(_HERE, "..", "include", "kdpc.h"))
    newest_header = max(os.path.getmtime(h) for h in headers if os.path.exists(h))

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            cmd = [nvcc] + COMPILE_FLAGS + ["-c", src, "-o", obj]
            log.append(" ".join(cmd))
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc] + LINK_FLAGS + ["-o", LIB_PATH] + objs
    log.append(" ".join(cmd))
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    with open(os.path.join(objdir, "build.log"), "w") as f:   # the nvcc command lines of the last (re)build
        f.write("\n".join(log) + "\n")
    return LIB_PATH


_P = c_void_p
_SIGNATURES = {
    # name: argtypes (restype is int unless noted)
    "kdpc_pointconv_fused_ordered": [c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float, _P, c_int,
                                     _P, _P, _P],
    "kdpc_spatial_sort_order_offset": [c_int],
    "kdpc_spatial_sort_order_stride": [c_int],
    "kdpc_flow_metrics": [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_fps": [c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_gather": [c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_gather_grad": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P],
    "kdpc_group": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_group_grad": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P],
    "kdpc_three_nn": [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P],
    "kdpc_three_interpolate": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P],
    "kdpc_three_interpolate_grad": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P],
    "kdpc_ball_query": [c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P],
    "kdpc_square_distance": [c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_knn": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_knn_bruteforce": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_spatial_sort": [c_int, c_int, _P, _P, _P],
    "kdpc_spatial_reorder": [c_int, c_int, _P, _P, _P, _P],
    "kdpc_dataprep_mask": [c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_dataprep_select": [c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_linear_dw": [c_longlong, c_int, c_int, _P, c_int, _P, c_int, _P, _P, c_int, _P, _P],
    "kdpc_weightnet_grad": [c_longlong, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_knn_feat": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P],
    "kdpc_knn_sorted": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P],
    "kdpc_gather_rows": [c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_group_concat": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P],
    "kdpc_concat_rows": [c_longlong, c_int, _P, _P, _P, _P, c_int, _P],
    "kdpc_weightnet": [c_longlong, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_pointconv_agg": [c_longlong, c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_pointconv_agg_grad": [c_longlong, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P],
    "kdpc_costvol_pre": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P],
    "kdpc_max_over_k": [c_longlong, c_int, c_int, _P, _P, _P, _P],
    "kdpc_interp3": [c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_pack_weight": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P],
    "kdpc_linear_tc": [c_longlong, c_int, c_int, _P, c_int, _P, _P, _P, c_float, c_float, c_float, _P, _P, _P, c_int, _P],
    "kdpc_linear_simt": [c_longlong, c_int, c_int, _P, c_int, _P, _P, _P, c_float, c_float, c_float, _P, _P, c_int, _P],
    "kdpc_pointconv_fused": [c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, _P],
    "kdpc_costvol_fused": [c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_float, _P, _P, _P],
    "kdpc_flow_loss": [c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P],
    "kdpc_hint_loss": [c_longlong, _P, _P, c_float, _P, _P, _P, _P],
    "kdpc_build_csr": [c_int, c_int, c_int, _P, _P, _P, _P],
    "kdpc_adam_step": [c_int, _P, _P, _P, _P, _P, _P, c_float, c_float, c_float, c_float, _P, _P],
    "kdpc_costvol_grad": [c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _P, _P, _P],
    "kdpc_scatter_rows_csr": [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P],
}

_lib = None


class KdpcError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libkdpc.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KdpcError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for the kdpc ops.")
        L = ctypes.CDLL(LIB_PATH)
        L.kdpc_abi_version.restype = c_int
        L.kdpc_error_string.restype = c_char_p
        L.kdpc_error_string.argtypes = [c_int]
        L.kdpc_packed_weight_bytes.restype = c_longlong
        L.kdpc_packed_weight_bytes.argtypes = [c_int, c_int]
        L.kdpc_knn_workspace_bytes.restype = c_longlong
        L.kdpc_knn_workspace_bytes.argtypes = [c_int, c_int, c_int]
        L.kdpc_spatial_sort_bytes.restype = c_longlong
        L.kdpc_spatial_sort_bytes.argtypes = [c_int, c_int]
        L.kdpc_linear_tc_ws_bytes.restype = c_longlong
        L.kdpc_linear_tc_ws_bytes.argtypes = [c_longlong, c_int, c_int]
        L.kdpc_pointconv_fused_ws_bytes.restype = c_longlong
        L.kdpc_pointconv_fused_ws_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
        L.kdpc_flow_metrics_workspace_bytes.restype = c_longlong
        L.kdpc_flow_metrics_workspace_bytes.argtypes = []
        L.kdpc_tc_set_pdl.restype = None
        L.kdpc_tc_set_pdl.argtypes = [c_int]
        if os.environ.get("KDPC_PDL", "0") == "1":               # A/B switch for measurements
            L.kdpc_tc_set_pdl(1)
        L.kdpc_knn_set_few.restype = None
        L.kdpc_knn_set_few.argtypes = [c_int]
        if os.environ.get("KDPC_KNN_FEW", "0") == "1":           # A/B switch for measurements (default off: slower)
            L.kdpc_knn_set_few(1)
        L.kdpc_group_concat_set_direct.restype = None
        L.kdpc_group_concat_set_direct.argtypes = [c_int]
        L.kdpc_costvol_grad_ws_bytes.restype = c_longlong
        L.kdpc_costvol_grad_ws_bytes.argtypes = []
        L.kdpc_costvol_fused_ws_bytes.restype = c_longlong
        L.kdpc_costvol_fused_ws_bytes.argtypes = [c_int, c_int, c_int, c_int]
        L.kdpc_dataprep_workspace_bytes.restype = c_longlong
        L.kdpc_dataprep_workspace_bytes.argtypes = [c_int, c_int]
        L.kdpc_linear_dw_ws_bytes.restype = c_longlong
        L.kdpc_linear_dw_ws_bytes.argtypes = [c_longlong, c_int, c_int]
        L.kdpc_weightnet_grad_ws_bytes.restype = c_longlong
        L.kdpc_weightnet_grad_ws_bytes.argtypes = [c_longlong]
        L.kdpc_loss_workspace_bytes.restype = c_longlong
        L.kdpc_loss_workspace_bytes.argtypes = []
        L.kdpc_set_sm_limit.restype = None
        L.kdpc_set_sm_limit.argtypes = [c_int]
        L.kdpc_sm_limit.restype = c_int
        L.kdpc_sm_limit.argtypes = []
        L.kdpc_fps_set_cluster.restype = None
        L.kdpc_fps_set_cluster.argtypes = [c_int]
        L.kdpc_fps_cluster_capacity.restype = c_int
        L.kdpc_fps_cluster_capacity.argtypes = [c_int, c_int, c_int]
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = c_int
        L.kdpc_tc_set_async.restype = None
        L.kdpc_tc_set_async.argtypes = [c_int]
        L.kdpc_tc_async_enabled.restype = c_int
        L.kdpc_pointconv_set_stages.restype = None
        L.kdpc_pointconv_set_stages.argtypes = [c_int]
        if os.environ.get("KDPC_PC_STAGES"):
            L.kdpc_pointconv_set_stages(int(os.environ["KDPC_PC_STAGES"]))
        L.kdpc_pointconv_set_precompute.restype = None
        L.kdpc_pointconv_set_precompute.argtypes = [c_int]
        if os.environ.get("KDPC_PC_PRECOMPUTE", "1") == "0":
            L.kdpc_pointconv_set_precompute(0)
        if os.environ.get("KDPC_TC_ASYNC", "2") != "2":          # A/B switch: 0 = synchronous producers, 1 = cp.async rows (2 = tensor-map TMA rows, default)
            L.kdpc_tc_set_async(int(os.environ["KDPC_TC_ASYNC"]))
        L.kdpc_linear_set_split_n.restype = None
        L.kdpc_linear_set_split_n.argtypes = [c_int]
        if os.environ.get("KDPC_SPLIT_N", "1") == "0":           # A/B switch for measurements
            L.kdpc_linear_set_split_n(0)
        L.kdpc_linear_dw_set_async.restype = None
        L.kdpc_linear_dw_set_async.argtypes = [c_int]
        if os.environ.get("KDPC_DW_ASYNC", "1") == "0":          # A/B switch for measurements
            L.kdpc_linear_dw_set_async(0)
        L.kdpc_costvol_set_pairing.restype = None
        L.kdpc_costvol_set_pairing.argtypes = [c_int]
        if os.environ.get("KDPC_CV_PAIR", "1") == "0":           # A/B switch for measurements
            L.kdpc_costvol_set_pairing(0)
        if os.environ.get("KDPC_FPS_CLUSTER", "1") != "1":       # A/B switch for measurements (0 = off, or a forced shape)
            L.kdpc_fps_set_cluster(int(os.environ["KDPC_FPS_CLUSTER"]))
        _lib = L
    return _lib


def exported_symbols():
    return ["kdpc_abi_version", "kdpc_error_string", "kdpc_set_sm_limit", "kdpc_sm_limit", "kdpc_packed_weight_bytes", "kdpc_knn_workspace_bytes",
            "kdpc_spatial_sort_bytes", "kdpc_costvol_fused_ws_bytes", "kdpc_linear_tc_ws_bytes", "kdpc_pointconv_fused_ws_bytes",
            "kdpc_loss_workspace_bytes", "kdpc_linear_dw_ws_bytes", "kdpc_weightnet_grad_ws_bytes", "kdpc_dataprep_workspace_bytes", "kdpc_flow_metrics_workspace_bytes", "kdpc_fps_set_cluster", "kdpc_fps_cluster_capacity", "kdpc_tc_set_async", "kdpc_tc_set_trace", "kdpc_tc_trace_buffer", "kdpc_pointconv_set_stages", "kdpc_pointconv_set_precompute",
            "kdpc_tc_async_enabled", "kdpc_costvol_set_pairing", "kdpc_linear_set_split_n", "kdpc_linear_dw_set_async", "kdpc_costvol_grad_ws_bytes", "kdpc_group_concat_set_direct", "kdpc_knn_set_few", "kdpc_tc_set_pdl", "kdpc_tc_pdl_enabled"] + list(_SIGNATURES)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().kdpc_error_string(rc).decode()
        raise KdpcError(f"{what} failed with status {rc}: {msg}")
