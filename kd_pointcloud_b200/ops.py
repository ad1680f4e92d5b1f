"""``torch.library`` registration of the libkdpc kernels as ``torch.ops.kdpc.*``.

Each op is registered for the CUDA dispatch key ONLY: a CPU tensor makes the dispatcher raise
("no kernel for the CPU backend") and a missing ``libkdpc.so`` raises ``KdpcError`` — there is no
fallback path.  The implementations are thin: validate, allocate the outputs/workspaces with the
torch allocator (the caller-allocates convention of the reference, pointnet2_utils.py:25-26), and
pass raw device pointers + the current stream through the C ABI (include/kdpc.h).

Differentiable wrappers live in ``functional.py``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch

from . import _lib
from ._lib import check

_LIB = torch.library.Library("kdpc", "DEF")
LAUNCHES = 0          # number of C-ABI calls issued (bench.py reports kernels launched)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, ndim=None, name="tensor"):
    if t.dtype != dtype:
        raise TypeError(f"kdpc: {name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"kdpc: {name} must be {ndim}-D, got shape {tuple(t.shape)}")
    # same contract as the reference wrappers (`assert xyz.is_contiguous()`, pointnet2_utils.py:22)
    assert t.is_contiguous(), f"kdpc: {name} must be contiguous"


TRACE = None          # set to a list to record (name, int args, start event, end event) per C-ABI call (tools/trace_model.py)


NVTX = os.environ.get("KDPC_NVTX", "0") == "1"     # one NVTX range per C-ABI call (visible in nsys / ncu --nvtx)


def _call(name: str, *args):
    global LAUNCHES
    LAUNCHES += 1
    if NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            check(getattr(_lib.lib(), name)(*args), name)
        finally:
            torch.cuda.nvtx.range_pop()
        return
    if TRACE is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(getattr(_lib.lib(), name)(*args), name)
        b.record()
        TRACE.append((name, tuple(x for x in args if isinstance(x, int) and not isinstance(x, bool) and abs(x) < (1 << 24)), a, b))
        return
    check(getattr(_lib.lib(), name)(*args), name)


class _guard:
    """Run on the device of the inputs (the reference uses the *current* device)."""

    def __init__(self, t: torch.Tensor):
        self.dev = t.device
        self.ctx = None

    def __enter__(self):
        if self.dev.index is not None and self.dev.index != torch.cuda.current_device():
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _register(schema: str, impl, fake):
    name = schema.split("(")[0]
    _LIB.define(schema)
    _LIB.impl(name, impl, "CUDA")
    torch.library.register_fake(f"kdpc::{name}")(fake)


# ------------------------------------------------------------------------------------------ a1
def _fps(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    _req(xyz, torch.float32, 3, "xyz")
    B, N, C = xyz.shape
    if C != 3:
        raise ValueError("kdpc: xyz must be [B,N,3]")
    with _guard(xyz):
        idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
        if B > 0 and npoint > 0:
            _call("kdpc_fps", B, N, npoint, _p(xyz), _p(temp), _p(idx), _stream())
    return idx


_register("fps(Tensor xyz, int npoint) -> Tensor", _fps,
          lambda xyz, npoint: xyz.new_empty((xyz.shape[0], npoint), dtype=torch.int32))


# ------------------------------------------------------------------------------------------ a2
def _gather_cm(f: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _req(f, torch.float32, 3, "features")
    _req(idx, torch.int32, 2, "idx")
    B, C, N = f.shape
    M = idx.shape[1]
    with _guard(f):
        out = torch.empty((B, C, M), dtype=torch.float32, device=f.device)
        if out.numel():
            _call("kdpc_gather", B, C, N, M, _p(f), _p(idx), _p(out), _stream())
    return out


def _gather_cm_grad(g: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _req(g, torch.float32, 3, "grad_out")
    _req(idx, torch.int32, 2, "idx")
    B, C, M = g.shape
    with _guard(g):
        gf = torch.empty((B, C, n), dtype=torch.float32, device=g.device)
        ws = torch.empty((B * (n + 1 + M),), dtype=torch.int32, device=g.device)
        _call("kdpc_gather_grad", B, C, n, M, _p(g), _p(idx), _p(ws), _p(gf), _stream())
    return gf


_register("gather_cm(Tensor f, Tensor idx) -> Tensor", _gather_cm,
          lambda f, idx: f.new_empty((f.shape[0], f.shape[1], idx.shape[1])))
_register("gather_cm_grad(Tensor g, Tensor idx, int n) -> Tensor", _gather_cm_grad,
          lambda g, idx, n: g.new_empty((g.shape[0], g.shape[1], n)))


# ------------------------------------------------------------------------------------------ a3
def _group_cm(f: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _req(f, torch.float32, 3, "features")
    _req(idx, torch.int32, 3, "idx")
    B, C, N = f.shape
    _, S, K = idx.shape
    with _guard(f):
        out = torch.empty((B, C, S, K), dtype=torch.float32, device=f.device)
        if out.numel():
            _call("kdpc_group", B, C, N, S, K, _p(f), _p(idx), _p(out), _stream())
    return out


def _group_cm_grad(g: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _req(g, torch.float32, 4, "grad_out")
    _req(idx, torch.int32, 3, "idx")
    B, C, S, K = g.shape
    with _guard(g):
        gf = torch.empty((B, C, n), dtype=torch.float32, device=g.device)
        ws = torch.empty((B * (n + 1 + S * K),), dtype=torch.int32, device=g.device)
        _call("kdpc_group_grad", B, C, n, S, K, _p(g), _p(idx), _p(ws), _p(gf), _stream())
    return gf


_register("group_cm(Tensor f, Tensor idx) -> Tensor", _group_cm,
          lambda f, idx: f.new_empty((f.shape[0], f.shape[1], idx.shape[1], idx.shape[2])))
_register("group_cm_grad(Tensor g, Tensor idx, int n) -> Tensor", _group_cm_grad,
          lambda g, idx, n: g.new_empty((g.shape[0], g.shape[1], n)))


# ------------------------------------------------------------------------------------- a4, a5
def _three_nn(unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(unknown, torch.float32, 3, "unknown")
    _req(known, torch.float32, 3, "known")
    B, N, _ = unknown.shape
    M = known.shape[1]
    with _guard(unknown):
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        ws = torch.empty((_lib.lib().kdpc_knn_workspace_bytes(B, N, M),), dtype=torch.uint8, device=unknown.device)
        if dist2.numel():
            _call("kdpc_three_nn", B, N, M, _p(unknown), _p(known), _p(ws), _p(dist2), _p(idx), _stream())
    return dist2, idx


def _three_interpolate_cm(f: torch.Tensor, idx: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    _req(f, torch.float32, 3, "features")
    _req(idx, torch.int32, 3, "idx")
    _req(w, torch.float32, 3, "weight")
    B, C, M = f.shape
    N = idx.shape[1]
    with _guard(f):
        out = torch.empty((B, C, N), dtype=torch.float32, device=f.device)
        if out.numel():
            _call("kdpc_three_interpolate", B, C, M, N, _p(f), _p(idx), _p(w), _p(out), _stream())
    return out


def _three_interpolate_cm_grad(g: torch.Tensor, idx: torch.Tensor, w: torch.Tensor, m: int) -> torch.Tensor:
    _req(g, torch.float32, 3, "grad_out")
    _req(idx, torch.int32, 3, "idx")
    _req(w, torch.float32, 3, "weight")
    B, C, N = g.shape
    with _guard(g):
        gf = torch.empty((B, C, m), dtype=torch.float32, device=g.device)
        ws = torch.empty((B * (m + 1 + 3 * N),), dtype=torch.int32, device=g.device)
        _call("kdpc_three_interpolate_grad", B, C, N, m, _p(g), _p(idx), _p(w), _p(ws), _p(gf), _stream())
    return gf


_register("three_nn(Tensor unknown, Tensor known) -> (Tensor, Tensor)", _three_nn,
          lambda u, k: (u.new_empty(u.shape), u.new_empty(u.shape, dtype=torch.int32)))
_register("three_interpolate_cm(Tensor f, Tensor idx, Tensor w) -> Tensor", _three_interpolate_cm,
          lambda f, idx, w: f.new_empty((f.shape[0], f.shape[1], idx.shape[1])))
_register("three_interpolate_cm_grad(Tensor g, Tensor idx, Tensor w, int m) -> Tensor",
          _three_interpolate_cm_grad, lambda g, idx, w, m: g.new_empty((g.shape[0], g.shape[1], m)))


def _ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    _req(xyz, torch.float32, 3, "xyz")
    _req(new_xyz, torch.float32, 3, "new_xyz")
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    with _guard(xyz):
        idx = torch.empty((B, M, nsample), dtype=torch.int32, device=xyz.device)
        if idx.numel():
            _call("kdpc_ball_query", B, N, M, float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _stream())
    return idx


_register("ball_query(float radius, int nsample, Tensor xyz, Tensor new_xyz) -> Tensor", _ball_query,
          lambda r, ns, xyz, new_xyz: xyz.new_empty((xyz.shape[0], new_xyz.shape[1], ns), dtype=torch.int32))


# ------------------------------------------------------------------------------------- a6, a7
def _square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    _req(src, torch.float32, 3, "src")
    _req(dst, torch.float32, 3, "dst")
    B, S, _ = src.shape
    N = dst.shape[1]
    with _guard(src):
        out = torch.empty((B, S, N), dtype=torch.float32, device=src.device)
        if out.numel():
            _call("kdpc_square_distance", B, S, N, _p(src), _p(dst), _p(out), _stream())
    return out


def _knn_impl(query, cand, k, want64, want_dist):
    _req(query, torch.float32, 3, "new_xyz")
    _req(cand, torch.float32, 3, "xyz")
    if query.shape[2] != 3 or cand.shape[2] != 3:
        raise ValueError("kdpc: knn expects [B,*,3] coordinates")
    B, S, _ = query.shape
    N = cand.shape[1]
    if k > N:
        raise RuntimeError(f"kdpc: knn k={k} exceeds the number of candidates {N}")  # torch.topk raises too
    dev = query.device
    with _guard(query):
        idx32 = torch.empty((B, S, k), dtype=torch.int32, device=dev)
        idx64 = torch.empty((B, S, k), dtype=torch.int64, device=dev) if want64 else None
        dist = torch.empty((B, S, k), dtype=torch.float32, device=dev) if want_dist else None
        ws = torch.empty((_lib.lib().kdpc_knn_workspace_bytes(B, S, N),), dtype=torch.uint8, device=dev)
        if idx32.numel():
            _call("kdpc_knn", B, S, N, k, _p(query), _p(cand), _p(ws), _p(idx32), _p(idx64), _p(dist), _stream())
    return idx32, idx64, dist


def _knn(query: torch.Tensor, cand: torch.Tensor, k: int) -> torch.Tensor:
    return _knn_impl(query, cand, k, False, False)[0]


def _knn64(query: torch.Tensor, cand: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    i32, i64, _ = _knn_impl(query, cand, k, True, False)
    return i32, i64


def _knn_dist(query: torch.Tensor, cand: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    i32, _, d = _knn_impl(query, cand, k, False, True)
    return i32, d


_register("square_distance(Tensor src, Tensor dst) -> Tensor", _square_distance,
          lambda s, d: s.new_empty((s.shape[0], s.shape[1], d.shape[1])))
_register("knn(Tensor query, Tensor cand, int k) -> Tensor", _knn,
          lambda q, c, k: q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32))
_register("knn64(Tensor query, Tensor cand, int k) -> (Tensor, Tensor)", _knn64,
          lambda q, c, k: (q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32),
                           q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int64)))
_register("knn_dist(Tensor query, Tensor cand, int k) -> (Tensor, Tensor)", _knn_dist,
          lambda q, c, k: (q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32),
                           q.new_empty((q.shape[0], q.shape[1], k))))


def _knn_feat(query: torch.Tensor, cand: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """query [B,S,C], cand [B,N,C] (any C <= 512) -> (idx int32 [B,S,k], dist [B,S,k]), ascending (distance, index)."""
    _req(query, torch.float32, 3, "new_xyz")
    _req(cand, torch.float32, 3, "xyz")
    B, S, C = query.shape
    N = cand.shape[1]
    if cand.shape[0] != B or cand.shape[2] != C:
        raise ValueError("kdpc: knn_feat query / candidate shapes do not match")
    if k > N:
        raise RuntimeError(f"kdpc: knn k={k} exceeds the number of candidates {N}")
    with _guard(query):
        idx = torch.empty((B, S, k), dtype=torch.int32, device=query.device)
        dist = torch.empty((B, S, k), dtype=torch.float32, device=query.device)
        if idx.numel():
            _call("kdpc_knn_feat", B, S, N, C, k, _p(query), _p(cand), _p(idx), _p(dist), _stream())
    return idx, dist


_register("knn_feat(Tensor query, Tensor cand, int k) -> (Tensor, Tensor)", _knn_feat,
          lambda q, c, k: (q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32),
                           q.new_empty((q.shape[0], q.shape[1], k))))


def _knn_bruteforce(query: torch.Tensor, cand: torch.Tensor, k: int) -> torch.Tensor:
    _req(query, torch.float32, 3, "new_xyz")
    _req(cand, torch.float32, 3, "xyz")
    B, S, _ = query.shape
    N = cand.shape[1]
    with _guard(query):
        idx32 = torch.empty((B, S, k), dtype=torch.int32, device=query.device)
        ws = torch.empty((B * N * 4,), dtype=torch.float32, device=query.device)
        _call("kdpc_knn_bruteforce", B, S, N, k, _p(query), _p(cand), _p(ws), _p(idx32), None, None, _stream())
    return idx32


SORT_MAX_N = 16384
SORT_MIN_N = 256


def _spatial_sort(xyz: torch.Tensor) -> torch.Tensor:
    """xyz [B,N,3] -> opaque uint8 buffer holding the Morton-sorted clouds (kdpc_spatial_sort)."""
    _req(xyz, torch.float32, 3, "xyz")
    B, N, _ = xyz.shape
    with _guard(xyz):
        out = torch.empty((_lib.lib().kdpc_spatial_sort_bytes(B, N),), dtype=torch.uint8, device=xyz.device)
        _call("kdpc_spatial_sort", B, N, _p(xyz), _p(out), _stream())
    return out


def _spatial_reorder(xyz: torch.Tensor, parent_sorted: torch.Tensor) -> torch.Tensor:
    """Sorted-cloud buffer of xyz [B,N,3] in the order of ``parent_sorted`` (the buffer of a cloud xyz is a displaced
    copy of): no second sort, boxes recomputed (kdpc_spatial_reorder)."""
    _req(xyz, torch.float32, 3, "xyz")
    _req(parent_sorted, torch.uint8, 1, "sorted parent")
    B, N, _ = xyz.shape
    if parent_sorted.numel() != _lib.lib().kdpc_spatial_sort_bytes(B, N):
        raise ValueError("kdpc: the parent's sorted-cloud buffer does not match xyz")
    with _guard(xyz):
        out = torch.empty_like(parent_sorted)
        _call("kdpc_spatial_reorder", B, N, _p(xyz), _p(parent_sorted), _p(out), _stream())
    return out


_register("spatial_reorder(Tensor xyz, Tensor parent_sorted) -> Tensor", _spatial_reorder, lambda xyz, p: torch.empty_like(p))


def _knn_sorted(qsorted: torch.Tensor, csorted: torch.Tensor, b: int, s: int, n: int, k: int) -> torch.Tensor:
    _req(qsorted, torch.uint8, 1, "sorted queries")
    _req(csorted, torch.uint8, 1, "sorted candidates")
    if k > n:
        raise RuntimeError(f"kdpc: knn k={k} exceeds the number of candidates {n}")
    L = _lib.lib()
    if qsorted.numel() != L.kdpc_spatial_sort_bytes(b, s) or csorted.numel() != L.kdpc_spatial_sort_bytes(b, n):
        raise ValueError("kdpc: sorted-cloud buffer does not match (b, s, n)")
    with _guard(qsorted):
        idx32 = torch.empty((b, s, k), dtype=torch.int32, device=qsorted.device)
        if idx32.numel():
            _call("kdpc_knn_sorted", b, s, n, k, 0, _p(qsorted), _p(csorted), _p(idx32), None, None, _stream())
    return idx32


_register("knn_bruteforce(Tensor query, Tensor cand, int k) -> Tensor", _knn_bruteforce,
          lambda q, c, k: q.new_empty((q.shape[0], q.shape[1], k), dtype=torch.int32))
_register("spatial_sort(Tensor xyz) -> Tensor", _spatial_sort, lambda xyz: xyz.new_empty((1,), dtype=torch.uint8))
_register("knn_sorted(Tensor qsorted, Tensor csorted, int b, int s, int n, int k) -> Tensor", _knn_sorted,
          lambda q, c, b, s, n, k: q.new_empty((b, s, k), dtype=torch.int32))


# ------------------------------------------------------------------------------------- a8, a9
def _gather_rows(f: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """f [B,N,C], idx int32 [B,...] -> [B,...,C]."""
    _req(f, torch.float32, 3, "points")
    _req(idx, torch.int32, None, "idx")
    B, N, C = f.shape
    if idx.shape[0] != B:
        raise ValueError("kdpc: idx batch size does not match points")
    M = idx.numel() // max(B, 1)
    with _guard(f):
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=f.device)
        if out.numel():
            _call("kdpc_gather_rows", B, N, M, C, _p(f), _p(idx), _p(out), _stream())
    return out


def _group_concat(cand_xyz: torch.Tensor, query_xyz: torch.Tensor, feats: Optional[torch.Tensor],
                  idx: torch.Tensor) -> torch.Tensor:
    _req(cand_xyz, torch.float32, 3, "xyz")
    _req(query_xyz, torch.float32, 3, "new_xyz")
    _req(idx, torch.int32, 3, "idx")
    B, N, _ = cand_xyz.shape
    _, S, K = idx.shape
    D = 0
    if feats is not None:
        _req(feats, torch.float32, 3, "points")
        D = feats.shape[2]
    with _guard(cand_xyz):
        out = torch.empty((B, S, K, 3 + D), dtype=torch.float32, device=cand_xyz.device)
        if out.numel():
            _call("kdpc_group_concat", B, N, S, K, D, _p(cand_xyz), _p(query_xyz), _p(feats), _p(idx), _p(out),
                  _stream())
    return out


_register("gather_rows(Tensor f, Tensor idx) -> Tensor", _gather_rows,
          lambda f, idx: f.new_empty(tuple(idx.shape) + (f.shape[2],)))
_register("group_concat(Tensor cand_xyz, Tensor query_xyz, Tensor? feats, Tensor idx) -> Tensor", _group_concat,
          lambda c, q, f, idx: c.new_empty(tuple(idx.shape) + (3 + (0 if f is None else f.shape[2]),)))


# ------------------------------------------------------------------------------------ a10, a11
def _weightnet(x: torch.Tensor, w1, b1, w2, b2, w3, b3) -> torch.Tensor:
    """x [..., W>=3] contiguous (the first 3 of every row are the localized xyz) -> [..., wout]."""
    _req(x, torch.float32, None, "localized_xyz")
    for t in (w1, b1, w2, b2, w3, b3):
        _req(t, torch.float32, None, "weightnet parameter")
    W = x.shape[-1]
    rows = x.numel() // W
    h1, h2, wout = w1.shape[0], w2.shape[0], w3.shape[0]
    with _guard(x):
        out = torch.empty(tuple(x.shape[:-1]) + (wout,), dtype=torch.float32, device=x.device)
        if out.numel():
            _call("kdpc_weightnet", rows, _p(x), W, h1, h2, wout, _p(w1), _p(b1), _p(w2), _p(b2), _p(w3), _p(b3),
                  _p(out), _stream())
    return out


def _weightnet_grad(x: torch.Tensor, g_out: torch.Tensor, w1, b1, w2, b2, w3, b3, want_input: bool):
    """Backward of kdpc::weightnet: x [..., C>=3] (first 3 columns = localized xyz), g_out [..., W] ->
    (gw1 [8,3], gb1 [8], gw2 [8,8], gb2 [8], gw3 [W,8], gb3 [W], g_x3 [..., 3] or empty)."""
    _req(x, torch.float32, None, "localized_xyz")
    _req(g_out, torch.float32, None, "grad_out")
    wout = g_out.shape[-1]
    rows = g_out.numel() // max(wout, 1)
    if x.numel() // x.shape[-1] != rows:
        raise ValueError("kdpc: weightnet_grad row counts differ")
    ps = [t.detach().contiguous().float() for t in (w1, b1, w2, b2, w3, b3)]
    dev = x.device
    with _guard(x):
        outs = [torch.empty_like(t) for t in ps]
        gx = torch.empty(tuple(x.shape[:-1]) + (3,), dtype=torch.float32, device=dev) if want_input else torch.empty(0, device=dev)
        if rows == 0:
            return tuple(o.zero_() for o in outs) + (gx,)
        ws = torch.empty((_lib.lib().kdpc_weightnet_grad_ws_bytes(rows),), dtype=torch.uint8, device=dev)
        _call("kdpc_weightnet_grad", rows, _p(x), x.shape[-1], wout, _p(g_out), *[_p(t) for t in ps], _p(ws),
              *[_p(o) for o in outs], _p(gx) if want_input else None, _stream())
    return tuple(outs) + (gx,)


_register("weightnet_grad(Tensor x, Tensor g_out, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor w3, Tensor b3, bool want_input)"
          " -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", _weightnet_grad,
          lambda x, g, w1, b1, w2, b2, w3, b3, wi: (torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2), torch.empty_like(b2),
                                                    torch.empty_like(w3), torch.empty_like(b3), x.new_empty(tuple(x.shape[:-1]) + (3,)) if wi else x.new_empty(0)))


def _pointconv_agg(grouped: torch.Tensor, wn: torch.Tensor) -> torch.Tensor:
    _req(grouped, torch.float32, 4, "grouped")
    _req(wn, torch.float32, 4, "weights")
    B, S, K, C = grouped.shape
    W = wn.shape[3]
    with _guard(grouped):
        out = torch.empty((B, S, C * W), dtype=torch.float32, device=grouped.device)
        if out.numel():
            _call("kdpc_pointconv_agg", B * S, K, C, W, _p(grouped), _p(wn), _p(out), _stream())
    return out


def _pointconv_agg_grad(grouped: torch.Tensor, wn: torch.Tensor, grad_out: torch.Tensor, want_grouped: bool, want_wn: bool):
    _req(grouped, torch.float32, 4, "grouped")
    _req(wn, torch.float32, 4, "weights")
    _req(grad_out, torch.float32, 3, "grad_out")
    B, S, K, C = grouped.shape
    W = wn.shape[3]
    if tuple(wn.shape[:3]) != (B, S, K) or tuple(grad_out.shape) != (B, S, C * W):
        raise ValueError("kdpc: pointconv_agg_grad shape mismatch")
    with _guard(grouped):
        gg = torch.empty_like(grouped) if want_grouped else torch.empty((0,), device=grouped.device)
        gw = torch.empty_like(wn) if want_wn else torch.empty((0,), device=grouped.device)
        if B * S and (want_grouped or want_wn):
            _call("kdpc_pointconv_agg_grad", B * S, K, C, W, _p(grouped), _p(wn), _p(grad_out),
                  _p(gg) if want_grouped else None, _p(gw) if want_wn else None, _stream())
    return gg, gw


_register("weightnet(Tensor x, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor w3, Tensor b3) -> Tensor",
          _weightnet, lambda x, w1, b1, w2, b2, w3, b3: x.new_empty(tuple(x.shape[:-1]) + (w3.shape[0],)))
_register("pointconv_agg_grad(Tensor grouped, Tensor wn, Tensor grad_out, bool want_grouped, bool want_wn) -> (Tensor, Tensor)",
          _pointconv_agg_grad,
          lambda g, w, go, a, b: (g.new_empty(g.shape if a else (0,)), w.new_empty(w.shape if b else (0,))))
_register("pointconv_agg(Tensor grouped, Tensor wn) -> Tensor", _pointconv_agg,
          lambda g, w: g.new_empty((g.shape[0], g.shape[1], g.shape[3] * w.shape[3])))


# ----------------------------------------------------------------------------------------- a13
def _costvol_pre(xyz1, xyz2, p1, p2, idx, pos_w, pos_b, slope: float) -> torch.Tensor:
    for t, nm in ((xyz1, "xyz1"), (xyz2, "xyz2"), (p1, "points1"), (p2, "points2")):
        _req(t, torch.float32, 3, nm)
    _req(idx, torch.int32, 3, "idx")
    _req(pos_w, torch.float32, None, "pos weight")
    _req(pos_b, torch.float32, 1, "pos bias")
    B, S, _ = xyz1.shape
    N = xyz2.shape[1]
    K = idx.shape[2]
    D = p1.shape[2]
    with _guard(xyz1):
        out = torch.empty((B, S, K, D), dtype=torch.float32, device=xyz1.device)
        if out.numel():
            _call("kdpc_costvol_pre", B, S, N, K, D, _p(xyz1), _p(xyz2), _p(p1), _p(p2), _p(idx), _p(pos_w),
                  _p(pos_b), float(slope), _p(out), _stream())
    return out


def _max_over_k(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(x, torch.float32, 4, "x")
    B, S, K, D = x.shape
    with _guard(x):
        out = torch.empty((B, S, D), dtype=torch.float32, device=x.device)
        arg = torch.empty((B, S, D), dtype=torch.int32, device=x.device)
        if out.numel():
            _call("kdpc_max_over_k", B * S, K, D, _p(x), _p(out), _p(arg), _stream())
    return out, arg


_register("costvol_pre(Tensor xyz1, Tensor xyz2, Tensor p1, Tensor p2, Tensor idx, Tensor pos_w, Tensor pos_b, "
          "float slope) -> Tensor", _costvol_pre,
          lambda x1, x2, p1, p2, idx, pw, pb, s: p1.new_empty(tuple(idx.shape) + (p1.shape[2],)))
_register("max_over_k(Tensor x) -> (Tensor, Tensor)", _max_over_k,
          lambda x: (x.new_empty((x.shape[0], x.shape[1], x.shape[3])),
                     x.new_empty((x.shape[0], x.shape[1], x.shape[3]), dtype=torch.int32)))


# ------------------------------------------------------------------------------------ a14, a15
def _interp3(q_xyz, c_xyz, idx, feat) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(q_xyz, torch.float32, 3, "xyz")
    _req(c_xyz, torch.float32, 3, "sparse_xyz")
    _req(idx, torch.int32, 3, "idx")
    _req(feat, torch.float32, 3, "sparse_flow")
    B, N, _ = q_xyz.shape
    S = c_xyz.shape[1]
    C = feat.shape[2]
    with _guard(q_xyz):
        out = torch.empty((B, N, C), dtype=torch.float32, device=q_xyz.device)
        w = torch.empty((B, N, 3), dtype=torch.float32, device=q_xyz.device)
        if out.numel():
            _call("kdpc_interp3", B, N, S, C, _p(q_xyz), _p(c_xyz), _p(idx), _p(feat), _p(out), _p(w), _stream())
    return out, w


_register("interp3(Tensor q_xyz, Tensor c_xyz, Tensor idx, Tensor feat) -> (Tensor, Tensor)", _interp3,
          lambda q, c, idx, f: (q.new_empty((q.shape[0], q.shape[1], f.shape[2])), q.new_empty(q.shape)))


# ------------------------------------------------------------------- tensor-core linear layers
def _pack_weight(w: torch.Tensor, mode: int, d: int, wn: int) -> torch.Tensor:
    _req(w, torch.float32, 2, "weight")
    n, k_src = w.shape
    k_packed = (d + 4) * wn if mode == 1 else k_src
    with _guard(w):
        nbytes = _lib.lib().kdpc_packed_weight_bytes(n, k_packed)
        out = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
        _call("kdpc_pack_weight", n, k_src, mode, d, wn, _p(w), _p(out), _stream())
    return out


def _linear_args(x, n, scale, shift, residual):
    _req(x, torch.float32, None, "x")
    k = x.shape[-1]
    m = x.numel() // k
    for t, nm in ((scale, "scale"), (shift, "shift")):
        if t is not None:
            _req(t, torch.float32, 1, nm)
            if t.numel() != n:
                raise ValueError(f"kdpc: {nm} must have {n} elements")
    if residual is not None:
        _req(residual, torch.float32, None, "residual")
        if residual.numel() != m * n:
            raise ValueError("kdpc: residual shape mismatch")
    return m, k


def _linear_tc(x, wpacked, n: int, scale, shift, slope: float, lo: float, hi: float, residual) -> torch.Tensor:
    m, k = _linear_args(x, n, scale, shift, residual)
    with _guard(x):
        out = torch.empty(tuple(x.shape[:-1]) + (n,), dtype=torch.float32, device=x.device)
        nws = _lib.lib().kdpc_linear_tc_ws_bytes(m, n, k)
        ws = torch.empty((nws,), dtype=torch.uint8, device=x.device) if nws else None
        if out.numel():
            _call("kdpc_linear_tc", m, n, k, _p(x), k, _p(wpacked), _p(scale), _p(shift), float(slope), float(lo),
                  float(hi), _p(residual), _p(ws), _p(out), n, _stream())
    return out


def row_stride(t: torch.Tensor) -> Optional[int]:
    """Elements between consecutive ROWS of ``t`` seen as a [rows, C] matrix (last axis unit-strided, every leading
    axis a whole multiple of the one after it), or None when ``t`` is not such a view.  A column block
    ``buf[..., c0:c0+C]`` of a contiguous tensor qualifies (stride = buf's width), and so does ``buf[:B]``."""
    if t.dim() == 0 or (t.shape[-1] > 1 and t.stride(-1) != 1):
        return None
    ld, span = None, None
    for size, stride in zip(reversed(t.shape[:-1]), reversed(t.stride()[:-1])):
        if size == 1:
            continue
        if ld is None:
            ld, span = stride, stride * size
        elif stride != span:
            return None
        else:
            span = stride * size
    return t.shape[-1] if ld is None else ld


def concat_rows(parts, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.cat(parts, dim=-1) of float32 activations in ONE vectorised copy kernel; every part may be a row-strided
    view (``row_stride``: a column block of a wider tensor), all with the same leading shape.  Channel counts % 4 == 0.
    (torch's cat drops to a scalar kernel when an input is a strided view: 35 us instead of ~13 for the level-0 tensor.)
    Not a dispatcher op: inference only."""
    import ctypes
    parts = list(parts)
    if not 1 <= len(parts) <= 4:
        raise ValueError("kdpc: concat_rows takes one to four tensors")
    lead = tuple(parts[0].shape[:-1])
    rows = 1
    for d in lead:
        rows *= d
    lds, widths = [], []
    for t in parts:
        if t.dtype != torch.float32 or tuple(t.shape[:-1]) != lead or not t.is_cuda:
            raise ValueError("kdpc: concat_rows needs float32 CUDA tensors with the same leading shape")
        ld = row_stride(t)
        if ld is None or ld % 4 or t.shape[-1] % 4 or t.data_ptr() % 16:
            raise ValueError("kdpc: concat_rows needs row-strided views with 16-byte aligned rows and widths % 4 == 0")
        lds.append(ld)
        widths.append(t.shape[-1])
    wtot = sum(widths)
    with _guard(parts[0]):
        if out is None:
            out = torch.empty(lead + (wtot,), dtype=torch.float32, device=parts[0].device)
        ldo = row_stride(out)
        if out.dtype != torch.float32 or tuple(out.shape) != lead + (wtot,) or ldo is None or ldo % 4 or out.data_ptr() % 16:
            raise ValueError("kdpc: concat_rows output must be a row-strided float32 [..., sum of widths] view")
        n = len(parts)
        src = (ctypes.c_void_p * n)(*[t.data_ptr() for t in parts])
        ld_a = (ctypes.c_int * n)(*lds)
        w_a = (ctypes.c_int * n)(*widths)
        if rows:
            _call("kdpc_concat_rows", rows, n, ctypes.cast(src, ctypes.c_void_p), ctypes.cast(ld_a, ctypes.c_void_p),
                  ctypes.cast(w_a, ctypes.c_void_p), _p(out), ldo, _stream())
    return out


def linear_tc_into(x: torch.Tensor, wpacked: torch.Tensor, n: int, out: torch.Tensor, scale=None, shift=None,
                   slope: float = 1.0, lo: float = 1.0, hi: float = 0.0, residual=None) -> torch.Tensor:
    """out[..., :n] = epilogue(x W^T) where BOTH ``x`` and ``out`` may be row-strided views (``row_stride``): a column
    block of a wider activation buffer, or a batch half of one.  The kernels take the row strides of the operand and of
    the result (ldx / ldo of kdpc_linear_tc), so a layer reads from and writes into the concatenated tensors its
    neighbours use - no torch.cat before or after.  Not a dispatcher op (it mutates ``out``): inference only."""
    if x.dtype != torch.float32 or out.dtype != torch.float32:
        raise TypeError("kdpc: linear_tc_into needs float32 tensors")
    k = x.shape[-1]
    ldx, ldo = row_stride(x), row_stride(out)
    m = x.numel() // max(k, 1)
    if ldx is None or ldo is None or out.shape[-1] != n or out.numel() != m * n or ldx < k or ldo < n:
        raise ValueError("kdpc: linear_tc_into needs row-strided [..., K] / [..., N] views with the same number of rows")
    if (ldx % 4) or (ldo % 4) or (x.data_ptr() % 16) or (out.data_ptr() % 16):
        raise ValueError("kdpc: linear_tc_into needs 16-byte aligned rows")
    for t, nm in ((scale, "scale"), (shift, "shift")):
        if t is not None and (t.dtype != torch.float32 or t.numel() != n or not t.is_contiguous()):
            raise ValueError(f"kdpc: {nm} must be a contiguous float32 [{n}]")
    if residual is not None and (residual.dtype != torch.float32 or not residual.is_contiguous() or residual.numel() != m * n
                                 or ldo != n):
        raise ValueError("kdpc: linear_tc_into takes a residual only with a contiguous output of the same shape")
    with _guard(x):
        nws = _lib.lib().kdpc_linear_tc_ws_bytes(m, n, k)
        ws = torch.empty((nws,), dtype=torch.uint8, device=x.device) if nws else None
        if m * n:
            _call("kdpc_linear_tc", m, n, k, _p(x), ldx, _p(wpacked), _p(scale), _p(shift), float(slope), float(lo),
                  float(hi), _p(residual), _p(ws), _p(out), ldo, _stream())
    return out


def _linear_dw(dy: torch.Tensor, x: torch.Tensor, want_bias: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """(dW [N,K], db [N]) = (dY^T X, column sums of dY) for dy [..., N], x [..., K] (same leading shape): the weight and
    bias gradients of y = x W^T + b in ONE tcgen05 launch (db rides along as an all-ones column of X)."""
    _req(dy, torch.float32, None, "grad_out")
    _req(x, torch.float32, None, "x")
    n, k = dy.shape[-1], x.shape[-1]
    m = dy.numel() // max(n, 1)
    if x.numel() // max(k, 1) != m:
        raise ValueError("kdpc: linear_dw row counts differ")
    with _guard(x):
        dw = torch.empty((n, k), dtype=torch.float32, device=x.device)
        db = torch.empty((n if want_bias else 0,), dtype=torch.float32, device=x.device)
        if m == 0:
            return dw.zero_(), db.zero_()
        ws = torch.empty((_lib.lib().kdpc_linear_dw_ws_bytes(m, n, k),), dtype=torch.uint8, device=x.device)
        _call("kdpc_linear_dw", m, n, k, _p(dy), n, _p(x), k, _p(ws), _p(dw), k, _p(db) if want_bias else None, _stream())
    return dw, db


_register("linear_dw(Tensor dy, Tensor x, bool want_bias) -> (Tensor, Tensor)", _linear_dw,
          lambda dy, x, wb: (dy.new_empty((dy.shape[-1], x.shape[-1])), dy.new_empty((dy.shape[-1] if wb else 0,))))


def _linear_simt(x, w, scale, shift, slope: float, lo: float, hi: float, residual) -> torch.Tensor:
    _req(w, torch.float32, 2, "weight")
    n = w.shape[0]
    m, k = _linear_args(x, n, scale, shift, residual)
    if w.shape[1] != k:
        raise ValueError("kdpc: weight / input size mismatch")
    with _guard(x):
        out = torch.empty(tuple(x.shape[:-1]) + (n,), dtype=torch.float32, device=x.device)
        if out.numel():
            _call("kdpc_linear_simt", m, n, k, _p(x), k, _p(w), _p(scale), _p(shift), float(slope), float(lo), float(hi),
                  _p(residual), _p(out), n, _stream())
    return out


_register("pack_weight(Tensor w, int mode, int d, int wn) -> Tensor", _pack_weight,
          lambda w, mode, d, wn: w.new_empty((1,), dtype=torch.uint8))
_register("linear_tc(Tensor x, Tensor wpacked, int n, Tensor? scale, Tensor? shift, float slope, float lo, float hi, "
          "Tensor? residual) -> Tensor", _linear_tc,
          lambda x, wp, n, sc, sh, sl, lo, hi, r: x.new_empty(tuple(x.shape[:-1]) + (n,)))
_register("linear_simt(Tensor x, Tensor w, Tensor? scale, Tensor? shift, float slope, float lo, float hi, "
          "Tensor? residual) -> Tensor", _linear_simt,
          lambda x, w, sc, sh, sl, lo, hi, r: x.new_empty(tuple(x.shape[:-1]) + (w.shape[0],)))


# ------------------------------------------------------------------- fused tcgen05 layers
def _pointconv_fused(cand_xyz, query_xyz, feats, idx, wn_params, wpacked, n_out: int, scale, shift,
                     slope: float, order=None) -> torch.Tensor:
    import ctypes
    _req(cand_xyz, torch.float32, 3, "xyz")
    _req(query_xyz, torch.float32, 3, "new_xyz")
    _req(feats, torch.float32, 3, "points")
    _req(idx, torch.int32, 3, "idx")
    B, N, _ = cand_xyz.shape
    _, S, K = idx.shape
    D = feats.shape[2]
    if tuple(query_xyz.shape[:2]) != (B, S) or tuple(feats.shape[:2]) != (B, N) or idx.shape[0] != B:
        raise ValueError(f"kdpc: pointconv_fused shape mismatch: xyz {tuple(cand_xyz.shape)}, new_xyz {tuple(query_xyz.shape)}, "
                         f"points {tuple(feats.shape)}, idx {tuple(idx.shape)}")
    if len(wn_params) != 248:
        raise ValueError("kdpc: pointconv_fused expects the 248 WeightNet(3,8,8,16) parameters")
    host = (ctypes.c_float * 248)(*wn_params)
    order_stride = 0
    if order is not None:          # int32 [B,S] rows of a (possibly wider) table: only the row stride may differ from S
        if order.dtype != torch.int32 or order.dim() != 2 or tuple(order.shape) != (B, S) or order.stride(1) != 1 \
                or order.device != feats.device or (B > 1 and order.stride(0) < S):
            raise ValueError("kdpc: pointconv_fused order must be int32 [B,S] with contiguous rows on the same device")
        order_stride = order.stride(0) if B > 1 else S
    with _guard(feats):
        out = torch.empty((B, S, n_out), dtype=torch.float32, device=feats.device)
        nws = _lib.lib().kdpc_pointconv_fused_ws_bytes(B, S, K, D, n_out)
        ws = torch.empty((nws,), dtype=torch.uint8, device=feats.device) if nws else None
        if out.numel():
            _call("kdpc_pointconv_fused_ordered", B, N, S, K, D, n_out, _p(cand_xyz), _p(query_xyz), _p(feats), _p(idx),
                  ctypes.cast(host, ctypes.c_void_p), _p(wpacked), _p(scale), _p(shift), float(slope),
                  None if order is None else order.data_ptr(), order_stride, _p(ws), _p(out), _stream())
    return out


def _costvol_fused(xyz1, xyz2, p1, p2, idx, pos_w, pos_b, slope_pre: float, wpacked, n_out: int, bias,
                   slope_post: float) -> torch.Tensor:
    for t, nm in ((xyz1, "xyz1"), (xyz2, "xyz2"), (p1, "points1"), (p2, "points2")):
        _req(t, torch.float32, 3, nm)
    _req(idx, torch.int32, 3, "idx")
    _req(pos_w, torch.float32, None, "pos weight")
    _req(pos_b, torch.float32, 1, "pos bias")
    B, S, _ = xyz1.shape
    N = xyz2.shape[1]
    K = idx.shape[2]
    D = p1.shape[2]
    with _guard(p1):
        out = torch.empty((B, S, n_out), dtype=torch.float32, device=p1.device)
        ws = torch.empty((_lib.lib().kdpc_costvol_fused_ws_bytes(B, S, N, D),), dtype=torch.uint8, device=p1.device)
        if out.numel():
            _call("kdpc_costvol_fused", B, S, N, K, D, n_out, _p(xyz1), _p(xyz2), _p(p1), _p(p2), _p(idx), _p(pos_w),
                  _p(pos_b), float(slope_pre), _p(wpacked), _p(bias), float(slope_post), _p(ws), _p(out), _stream())
    return out


def _costvol_grad(p1q, p2q, idx, w, bias, slope_pre: float, slope_post: float, grad_out):
    """(grad_p1q [B,S,D], grad_rows [B,S,K,D], grad_w [D',D], grad_b [D']) of the fused cost volume in its folded form."""
    for t, nm in ((p1q, "p1q"), (p2q, "p2q"), (grad_out, "grad_out")):
        _req(t, torch.float32, 3, nm)
    _req(idx, torch.int32, 3, "idx")
    _req(w, torch.float32, 2, "weight")
    B, S, D = p1q.shape
    N, K, Dout = p2q.shape[1], idx.shape[2], w.shape[0]
    if grad_out.shape != (B, S, Dout) or w.shape[1] != D or p2q.shape[2] != D:
        raise ValueError("kdpc: costvol_grad shape mismatch")
    with _guard(p1q):
        g1 = torch.empty_like(p1q)
        grows = torch.empty((B, S, K, D), dtype=torch.float32, device=p1q.device)
        gw = torch.empty_like(w)
        gb = torch.empty((Dout,), dtype=torch.float32, device=p1q.device)
        ws = torch.empty((_lib.lib().kdpc_costvol_grad_ws_bytes(),), dtype=torch.uint8, device=p1q.device)
        _call("kdpc_costvol_grad", B, S, N, K, D, Dout, _p(p1q), _p(p2q), _p(idx), _p(w), _p(bias), float(slope_pre),
              float(slope_post), _p(grad_out), _p(ws), _p(g1), _p(grows), _p(gw), _p(gb), _stream())
    return g1, grows, gw, gb


_register("costvol_grad(Tensor p1q, Tensor p2q, Tensor idx, Tensor w, Tensor? bias, float slope_pre, float slope_post, "
          "Tensor grad_out) -> (Tensor, Tensor, Tensor, Tensor)", _costvol_grad,
          lambda p1q, p2q, idx, w, b, s0, s1, g: (torch.empty_like(p1q), p1q.new_empty(tuple(idx.shape) + (p1q.shape[2],)),
                                                  torch.empty_like(w), w.new_empty((w.shape[0],))))
_register("pointconv_fused(Tensor cand_xyz, Tensor query_xyz, Tensor feats, Tensor idx, float[] wn_params, "
          "Tensor wpacked, int n_out, Tensor? scale, Tensor? shift, float slope, Tensor? order=None) -> Tensor",
          _pointconv_fused,
          lambda c, q, f, idx, wn, wp, n, sc, sh, sl, order=None: f.new_empty((idx.shape[0], idx.shape[1], n)))
_register("costvol_fused(Tensor xyz1, Tensor xyz2, Tensor p1, Tensor p2, Tensor idx, Tensor pos_w, Tensor pos_b, "
          "float slope_pre, Tensor wpacked, int n_out, Tensor? bias, float slope_post) -> Tensor", _costvol_fused,
          lambda x1, x2, p1, p2, idx, pw, pb, s0, wp, n, b, s1: p1.new_empty((p1.shape[0], p1.shape[1], n)))


# ------------------------------------------------------------------------------------ a18, a19
_LOSS_WS = {}


def _loss_ws(device) -> torch.Tensor:
    ws = _LOSS_WS.get(device)
    if ws is None:
        ws = torch.zeros((_lib.lib().kdpc_loss_workspace_bytes(),), dtype=torch.uint8, device=device)
        _LOSS_WS[device] = ws
    return ws


def _kd_loss(preds, fps_idxs, targets, alpha, weights, point_major: bool, hint_fs, hint_ft, hint_w, want_grad: bool):
    """loss[1] = sum_t weights[t] * multiScale(preds, targets[t]) + sum_h 0.5 * hint_w[h] * |hint_fs[h] - hint_ft[h]|^2,
    plus (want_grad) the gradients w.r.t. preds and hint_fs.  One flow kernel + one kernel per hint pair, all
    accumulating into the same scalar with deterministic reductions."""
    import ctypes
    ns = len(preds)
    if not (1 <= ns <= 4) or len(fps_idxs) < ns - 1 or not (1 <= len(targets) <= 2) or len(alpha) < ns:
        raise ValueError("kdpc: kd_loss supports 1..4 scales, 1..2 targets")
    if len(weights) != len(targets) or len(hint_fs) != len(hint_ft) or len(hint_fs) != len(hint_w):
        raise ValueError("kdpc: kd_loss list lengths do not match")
    B = preds[0].shape[0]
    n = []
    for s_, pr in enumerate(preds):
        _req(pr, torch.float32, 3, "pred_flow")
        if pr.shape[0] != B or pr.shape[2 if point_major else 1] != 3:
            raise ValueError("kdpc: pred_flow must be [B,3,N] (or [B,N,3] when point_major)")
        n.append(pr.shape[1 if point_major else 2])
    fps_idxs = list(fps_idxs)[len(fps_idxs) - (ns - 1):] if ns > 1 else []
    for s_, ix in enumerate(fps_idxs):
        _req(ix, torch.int32, 2, "fps_idx")
        if tuple(ix.shape) != (B, n[s_ + 1]):
            raise ValueError("kdpc: fps_idx[i] must be [B, N_{i+1}]")
    for t in targets:
        _req(t, torch.float32, 3, "target flow")
        if tuple(t.shape) != (B, n[0], 3):
            raise ValueError("kdpc: target flow must be [B,N0,3]")
    dev = preds[0].device
    with _guard(preds[0]):
        loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        grads = [torch.empty_like(pr) for pr in preds] if want_grad else []
        ws = _loss_ws(dev)
        ptr_arr = lambda ts: (ctypes.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts])
        n_arr = (ctypes.c_int * ns)(*n)
        alpha_arr = (ctypes.c_float * ns)(*[float(a) for a in alpha[:ns]])
        w_arr = (ctypes.c_float * len(targets))(*[float(w) for w in weights])
        _call("kdpc_flow_loss", B, ns, 1 if point_major else 0, ctypes.cast(n_arr, ctypes.c_void_p),
              ctypes.cast(ptr_arr(preds), ctypes.c_void_p),
              ctypes.cast(ptr_arr(grads), ctypes.c_void_p) if want_grad else None,
              ctypes.cast(ptr_arr(fps_idxs), ctypes.c_void_p), ctypes.cast(alpha_arr, ctypes.c_void_p), len(targets),
              ctypes.cast(ptr_arr(targets), ctypes.c_void_p), ctypes.cast(w_arr, ctypes.c_void_p), _p(ws), _p(loss),
              _stream())
        hgrads = []
        for fs, ft, w in zip(hint_fs, hint_ft, hint_w):
            _req(fs, torch.float32, None, "student feature")
            _req(ft, torch.float32, None, "teacher feature")
            if fs.shape != ft.shape:
                raise RuntimeError(f"kdpc: hint features differ in shape: {tuple(fs.shape)} vs {tuple(ft.shape)}")
            g = torch.empty_like(fs) if want_grad else None
            if fs.numel():
                _call("kdpc_hint_loss", fs.numel(), _p(fs), _p(ft), float(w), _p(g), _p(ws), _p(loss), _stream())
            if want_grad:
                hgrads.append(g)
    return loss, grads, hgrads


_register("kd_loss(Tensor[] preds, Tensor[] fps_idxs, Tensor[] targets, float[] alpha, float[] weights, bool point_major, "
          "Tensor[] hint_fs, Tensor[] hint_ft, float[] hint_w, bool want_grad) -> (Tensor, Tensor[], Tensor[])", _kd_loss,
          lambda preds, fps, tg, al, w, pm, hs, ht, hw, wg: (preds[0].new_empty((1,)),
                                                              [torch.empty_like(p) for p in preds] if wg else [],
                                                              [torch.empty_like(h) for h in hs] if wg else []))


# --------------------------------------------------------------------------- backward plumbing
def _build_csr(idx: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _req(idx, torch.int32, None, "idx")
    B = idx.shape[0]
    M = idx.numel() // B
    with _guard(idx):
        offsets = torch.empty((B, n + 1), dtype=torch.int32, device=idx.device)
        perm = torch.empty((B, M), dtype=torch.int32, device=idx.device)
        _call("kdpc_build_csr", B, n, M, _p(idx), _p(offsets), _p(perm), _stream())
    return offsets, perm


def _scatter_rows_csr(g: torch.Tensor, wgt: Optional[torch.Tensor], offsets: torch.Tensor, perm: torch.Tensor,
                      n: int, gdiv: int) -> torch.Tensor:
    """g [B, M/gdiv, C] (any leading layout flattening to that) -> [B, n, C]."""
    _req(g, torch.float32, None, "grad")
    B = offsets.shape[0]
    M = perm.shape[1]
    C = g.shape[-1]
    if g.numel() != B * (M // gdiv) * C:
        raise ValueError("kdpc: scatter_rows_csr shape mismatch")
    if wgt is not None:
        _req(wgt, torch.float32, None, "weight")
    with _guard(g):
        out = torch.empty((B, n, C), dtype=torch.float32, device=g.device)
        _call("kdpc_scatter_rows_csr", B, n, M, C, gdiv, _p(g), _p(wgt), _p(offsets), _p(perm), _p(out), 0, _stream())
    return out


_register("build_csr(Tensor idx, int n) -> (Tensor, Tensor)", _build_csr,
          lambda idx, n: (idx.new_empty((idx.shape[0], n + 1)), idx.new_empty((idx.shape[0], idx.numel() // idx.shape[0]))))
_register("scatter_rows_csr(Tensor g, Tensor? wgt, Tensor offsets, Tensor perm, int n, int gdiv) -> Tensor",
          _scatter_rows_csr, lambda g, w, o, p, n, gdiv: g.new_empty((o.shape[0], n, g.shape[-1])))

kdpc = torch.ops.kdpc


# ------------------------------------------------------------------------------------ evaluation metrics (8f-2)
_METRIC_WS = {}


def _flow_metrics(pred: torch.Tensor, gt: torch.Tensor, pc1, calib, point_major: bool) -> torch.Tensor:
    """float32 [6] on the device: EPE3D, Acc3DS, Acc3DR, Outliers3D, EPE2D, Acc2D of one batch (include/kdpc.h)."""
    _req(pred, torch.float32, 3, "pred_flow")
    _req(gt, torch.float32, 3, "gt_flow")
    B, N, C = gt.shape
    if C != 3 or tuple(pred.shape) != ((B, N, 3) if point_major else (B, 3, N)):
        raise ValueError("kdpc: flow_metrics expects gt [B,N,3] and pred [B,3,N] (or [B,N,3] when point_major)")
    if pc1 is not None:
        _req(pc1, torch.float32, 3, "pc1")
        if tuple(pc1.shape) != (B, N, 3):
            raise ValueError("kdpc: flow_metrics pc1 must be [B,N,3]")
    if calib is not None:
        _req(calib, torch.float32, 2, "calib")
        if tuple(calib.shape) != (B, 6):
            raise ValueError("kdpc: flow_metrics calib must be [B,6] (f, cx, cy, constx, consty, constz)")
    dev = gt.device
    with _guard(gt):
        ws = _METRIC_WS.get(dev)
        if ws is None:
            ws = _METRIC_WS[dev] = torch.zeros(((_lib.lib().kdpc_flow_metrics_workspace_bytes() + 7) // 8,),
                                               dtype=torch.int64, device=dev)
        out = torch.zeros((6,), dtype=torch.float32, device=dev)
        if B * N > 0:
            _call("kdpc_flow_metrics", B, N, 1 if point_major else 0, _p(pred), _p(gt), _p(pc1), _p(calib), _p(ws), _p(out),
                  _stream())
    return out


_register("flow_metrics(Tensor pred, Tensor gt, Tensor? pc1, Tensor? calib, bool point_major) -> Tensor", _flow_metrics,
          lambda pred, gt, pc1, calib, pm: gt.new_empty((6,)))
