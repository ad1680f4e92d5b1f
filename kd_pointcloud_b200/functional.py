"""Differentiable, cached front-end over ``torch.ops.kdpc``.

* Indices are constants for autograd (the reference returns ``None`` for them,
  pointnet2_utils.py:31-33,100-102); gradients flow to features and — where the reference's
  autograd would — to coordinates.
* Every gather-type backward is a fixed-order segmented reduction over an inverse index (CSR)
  instead of the reference's fp32 ``atomicAdd`` (sampling_gpu.cu:62, group_points_gpu.cu:24,
  interpolate_gpu.cu:139-141): deterministic, and the CSR is cached per index tensor so layers
  that share an index (the two flow-estimator PointConvs, cross call 1 and 3, the upsample calls
  of one level) build it once.
* ``knn_idx`` caches its result per (query, candidates, K) tensor identity + version: the
  reference recomputes 45 kNNs per forward of which only 29 are distinct.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import os
import torch

from . import _lib
from . import ops  # noqa: F401  (registers torch.ops.kdpc)

K = torch.ops.kdpc

# ------------------------------------------------------------------------------------ caches


def _tkey(t: torch.Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape), tuple(t.stride()), t.device.index)


# Weights can change WITHOUT their ``_version`` moving: a CUDA-graph replay of a training step (training.GraphedKDStep)
# updates parameters and BatchNorm statistics on the device only.  Everything derived from weights (packed bf16 images,
# folded BN affines, WeightNet host parameters) is therefore also keyed by this epoch, which graph owners bump after
# every replay (bump_weight_epoch): a later eager / eval forward re-derives instead of hitting a stale entry.
_WEIGHT_EPOCH = 0


def bump_weight_epoch() -> int:
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1
    return _WEIGHT_EPOCH


def _wkey(t: torch.Tensor):
    return _tkey(t) + (_WEIGHT_EPOCH,)


def weight_cache_tensors():
    """Every device tensor the weight-derived caches currently own (graph owners keep these alive: a captured graph
    bakes in the addresses of cache hits, and an LRU eviction or clear_caches(weights=True) must not free them)."""
    out = []
    for cache in (_PACK_CACHE, _AFFINE_CACHE):
        for v in cache.d.values():
            out.extend(t for t in v if isinstance(t, torch.Tensor))
    return out


class _LRU:
    def __init__(self, cap: int):
        self.cap = cap
        self.d: "OrderedDict" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get(self, key):
        v = self.d.get(key)
        if v is not None:
            self.d.move_to_end(key)
            self.hits += 1
        else:
            self.misses += 1
        return v

    def put(self, key, value):
        self.d[key] = value
        if len(self.d) > self.cap:
            self.d.popitem(last=False)

    def clear(self):
        self.d.clear()


_KNN_CACHE = _LRU(48)
_CSR_CACHE = _LRU(48)
_CACHE_ENABLED = True


_EXTRA_CACHES = []          # per-batch caches owned by other modules (flownet's sampling pyramid): cleared with the rest


def register_batch_cache(cache) -> None:
    _EXTRA_CACHES.append(cache)


def clear_caches(weights: bool = False) -> None:
    """Drop cached kNN results and inverse indices (call between independent batches);
    ``weights=True`` also drops the packed-weight / folded-affine caches."""
    _KNN_CACHE.clear()
    _CSR_CACHE.clear()
    _SORT_CACHE.clear()
    _SORT_PARENT.clear()
    for c in _EXTRA_CACHES:
        c.clear()
    if weights:
        _PACK_CACHE.clear()
        _AFFINE_CACHE.clear()
        _WN_CACHE.clear()


def snapshot_neighbour_caches():
    """The current kNN results and sorted clouds (what a coordinate-only pre-pass computed), to be re-installed later by
    ``seed_neighbour_caches`` - the runner computes them for batch i+1 on a second stream."""
    return dict(_KNN_CACHE.d), dict(_SORT_CACHE.d)


def seed_neighbour_caches(snapshot) -> None:
    knn, srt = snapshot
    for k, v in srt.items():
        _SORT_CACHE.put(k, v)
    for k, v in knn.items():
        _KNN_CACHE.put(k, v)


def set_cache_enabled(flag: bool) -> None:
    global _CACHE_ENABLED
    _CACHE_ENABLED = bool(flag)
    clear_caches()


def cache_stats():
    return {"knn_hits": _KNN_CACHE.hits, "knn_misses": _KNN_CACHE.misses,
            "csr_hits": _CSR_CACHE.hits, "csr_misses": _CSR_CACHE.misses}


def pm(x: torch.Tensor) -> torch.Tensor:
    """[B,C,N] -> contiguous point-major [B,N,C] (free when x is a permuted pm tensor)."""
    return x.permute(0, 2, 1).contiguous()


def cm(x: torch.Tensor) -> torch.Tensor:
    """point-major [B,N,C] -> the reference's [B,C,N] as a permuted view."""
    return x.permute(0, 2, 1)


# --------------------------------------------------------------------------- kNN (a6, a7)
_SORT_CACHE = _LRU(48)


_SORT_PARENT = _LRU(16)         # displaced cloud -> the cloud it was displaced from (PointWarping)
USE_SORT_REUSE = os.environ.get("KDPC_SORT_REUSE", "1") != "0"


def concat_free(*ts: torch.Tensor) -> bool:
    """True on the inference path (no autograd, float32 CUDA tensors): layers may then write into / read from column
    blocks and batch halves of shared activation buffers (``fused_linear(out=...)``) instead of producing tensors that
    are concatenated afterwards.  Same kernels, same values - only the destinations differ."""
    return CONCAT_FREE and not torch.is_grad_enabled() and all(t.is_cuda and t.dtype == torch.float32 for t in ts)


CONCAT_FREE = os.environ.get("KDPC_CONCAT_FREE", "1") != "0"      # A/B switch (tests compare both settings bit for bit)


def joined(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=0) - without the copy when a and b already ARE the two halves of one contiguous tensor
    (CrossLayerLight writes its two directions into one [2B,N,C] buffer on the inference path)."""
    base = a._base
    if (base is not None and b._base is base and base.is_contiguous() and a.is_contiguous() and b.is_contiguous()
            and base.dim() == a.dim() == b.dim() and base.shape[0] == a.shape[0] + b.shape[0]
            and tuple(base.shape[1:]) == tuple(a.shape[1:]) == tuple(b.shape[1:]) and a.data_ptr() == base.data_ptr()
            and b.data_ptr() == base.data_ptr() + a.numel() * a.element_size()):
        return base
    return torch.cat([a, b], dim=0)


def hint_displaced_copy(child: torch.Tensor, parent: torch.Tensor) -> None:
    """``child`` [B,N,3] is a smooth displacement of ``parent`` (same shape): when a kNN needs the child sorted, the parent's
    Morton order is reused (kdpc_spatial_reorder) instead of sorting again.  Results do not change."""
    if USE_SORT_REUSE and _CACHE_ENABLED and child.shape == parent.shape and child.is_cuda:
        c = child.detach()
        _SORT_PARENT.put(_tkey(c if c.is_contiguous() else c.contiguous()), (parent.detach(), child))


def _sorted_cloud(xyz_d: torch.Tensor) -> torch.Tensor:
    """Morton-sorted copy of a contiguous [B,N,3] cloud, cached per tensor (storage, version): every point
    set of the pyramid takes part in several kNN calls (as queries and as candidates)."""
    if not _CACHE_ENABLED:
        return K.spatial_sort(xyz_d)
    key = _tkey(xyz_d)
    hit = _SORT_CACHE.get(key)
    if hit is not None:
        return hit[0]
    # a batch slice of an already sorted batch (pc1 / pc2 halves of the 2B-cloud encoder batch): clouds are
    # sorted independently and stored back to back, so the slice of the buffer is the sorted slice
    B, N, _ = xyz_d.shape
    per = N * 12
    for pkey, (buf, parent) in list(_SORT_CACHE.d.items()):
        if parent.shape[1] != N or parent._version != pkey[1] or parent.device != xyz_d.device:
            continue
        off = xyz_d.data_ptr() - parent.data_ptr()
        if off >= 0 and off % per == 0 and off // per + B <= parent.shape[0]:
            each = buf.numel() // parent.shape[0]
            out = buf[(off // per) * each:(off // per + B) * each]
            _SORT_CACHE.put(key, (out, xyz_d))
            return out
    par = _SORT_PARENT.get(key)
    if par is not None and ops.SORT_MIN_N <= N <= ops.SORT_MAX_N:
        pd = par[0] if par[0].is_contiguous() else par[0].contiguous()
        out = K.spatial_reorder(xyz_d, _sorted_cloud(pd))          # the parent's order, boxes from the new coordinates
    else:
        out = K.spatial_sort(xyz_d)
    _SORT_CACHE.put(key, (out, xyz_d))
    return out


def _knn_compute(nsample: int, xyz_d: torch.Tensor, new_d: torch.Tensor) -> torch.Tensor:
    n, s = xyz_d.shape[1], new_d.shape[1]
    if xyz_d.shape[2] != 3:                        # feature-space kNN (CrossLayerLightFG, pointconv_util.py:1905)
        return K.knn_feat(new_d, xyz_d, nsample)[0]
    if ops.SORT_MIN_N <= n <= ops.SORT_MAX_N and s <= ops.SORT_MAX_N and xyz_d.shape[0] > 0:
        cs = _sorted_cloud(xyz_d)
        same = new_d.data_ptr() == xyz_d.data_ptr() and new_d.shape == xyz_d.shape
        qs = cs if same else _sorted_cloud(new_d)
        return K.knn_sorted(qs, cs, xyz_d.shape[0], s, n, nsample)
    return K.knn(new_d, xyz_d, nsample)


def knn_idx(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """int32 [B,S,nsample] neighbours of new_xyz in xyz, ascending (distance, index)."""
    xyz_d = xyz.detach()
    new_d = new_xyz.detach()
    if not xyz_d.is_contiguous():
        xyz_d = xyz_d.contiguous()
    if not new_d.is_contiguous():
        new_d = new_d.contiguous()
    if not _CACHE_ENABLED:
        return _knn_compute(nsample, xyz_d, new_d)
    key = (nsample, _tkey(xyz_d), _tkey(new_d))
    hit = _KNN_CACHE.get(key)
    if hit is not None:
        return hit[0]
    idx = _knn_compute(nsample, xyz_d, new_d)
    _KNN_CACHE.put(key, (idx, xyz_d, new_d))       # keep the keyed tensors alive: no address reuse
    return idx


def knn_point(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """pointconv_util.knn_point (pointconv_util.py:96-107): int64 [B,S,nsample]."""
    return knn_idx(nsample, xyz, new_xyz).long()


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """pointconv_util.square_distance (pointconv_util.py:73-94): [B,S,3], [B,N,3] -> [B,S,N], the kernel's values (the
    reference's rounding).  Differentiable like the reference's torch expression when an input requires a gradient
    (the self-supervised losses of models_bid_pointconv.py:574-638 use the distances by value): the gradient is that
    of the same expansion, attached to the kernel's values."""
    hard = K.square_distance(src.detach().contiguous(), dst.detach().contiguous())
    if not (torch.is_grad_enabled() and (src.requires_grad or dst.requires_grad)):
        return hard
    soft = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    soft = soft + torch.sum(src ** 2, -1).unsqueeze(2) + torch.sum(dst ** 2, -1).unsqueeze(1)
    return soft + (hard - soft).detach()


def _csr(idx: torch.Tensor, n: int):
    if not _CACHE_ENABLED:
        return K.build_csr(idx, n)
    key = (n, _tkey(idx))
    hit = _CSR_CACHE.get(key)
    if hit is not None:
        return hit[0], hit[1]
    off, perm = K.build_csr(idx, n)
    _CSR_CACHE.put(key, (off, perm, idx))
    return off, perm


def _as_i32(idx: torch.Tensor) -> torch.Tensor:
    if idx.dtype != torch.int32:
        idx = idx.int()                            # the reference's `.int()` (pointconv_util.py:131)
    return idx.contiguous()


# --------------------------------------------------------------- channel-major pointnet2 ops
class _GatherCM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return K.gather_cm(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return K.gather_cm_grad(grad_out.contiguous(), idx, ctx.n), None


class _GroupCM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return K.group_cm(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return K.group_cm_grad(grad_out.contiguous(), idx, ctx.n), None


class _ThreeInterpolateCM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.shape[2]
        return K.three_interpolate_cm(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return K.three_interpolate_cm_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    return K.fps(xyz.detach(), int(npoint))


def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return _GatherCM.apply(features, idx)


def grouping_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    return _GroupCM.apply(features, idx)


def three_nn(unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    dist2, idx = K.three_nn(unknown.detach(), known.detach())
    return torch.sqrt(dist2), idx                  # pointnet2_utils.py:98


def three_interpolate(features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    return _ThreeInterpolateCM.apply(features, idx, weight)


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    return K.ball_query(float(radius), int(nsample), xyz.detach(), new_xyz.detach())


# ------------------------------------------------------------------- point-major gathers
class _GatherRows(torch.autograd.Function):
    """out[b, ..., :] = points[b, idx[b, ...], :]"""

    @staticmethod
    def forward(ctx, points, idx):
        ctx.save_for_backward(idx)
        ctx.n = points.shape[1]
        return K.gather_rows(points, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        off, perm = _csr(idx, ctx.n)
        return K.scatter_rows_csr(grad_out.contiguous(), None, off, perm, ctx.n, 1), None


def gather_rows(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """points [B,N,C], idx [B,...] -> [B,...,C]  (index_points_gather / index_points_group, pm)."""
    return _GatherRows.apply(points.contiguous(), _as_i32(idx))


class _GroupConcat(torch.autograd.Function):
    """[cand_xyz[idx] - query_xyz (3) | feats[idx] (D)]  ->  [B,S,K,3+D]"""

    @staticmethod
    def forward(ctx, cand_xyz, query_xyz, feats, idx):
        ctx.save_for_backward(idx)
        ctx.n = cand_xyz.shape[1]
        ctx.has_feats = feats is not None
        ctx.same_xyz = cand_xyz.data_ptr() == query_xyz.data_ptr() and cand_xyz.shape == query_xyz.shape
        return K.group_concat(cand_xyz, query_xyz, feats, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        need_c, need_q, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g_c = g_q = g_f = None
        off = perm = None
        if need_f and ctx.has_feats:
            off, perm = _csr(idx, ctx.n)
            g_f = K.scatter_rows_csr(grad_out[..., 3:].contiguous(), None, off, perm, ctx.n, 1)
        if need_c or need_q:
            g_rel = grad_out[..., :3].contiguous()
            if need_c:
                if off is None:
                    off, perm = _csr(idx, ctx.n)
                g_c = K.scatter_rows_csr(g_rel, None, off, perm, ctx.n, 1)
            if need_q:
                g_q = -g_rel.sum(dim=2)
        return g_c, g_q, g_f, None


def group_concat(cand_xyz, query_xyz, feats: Optional[torch.Tensor], idx) -> torch.Tensor:
    return _GroupConcat.apply(cand_xyz.contiguous(), query_xyz.contiguous(),
                              None if feats is None else feats.contiguous(), _as_i32(idx))


# ------------------------------------------------------------------- 3-NN interpolation
class _Interp3(torch.autograd.Function):
    """Inverse-distance interpolation; differentiable w.r.t. the interpolated features only
    (coordinates never require grad in UpsampleFlow)."""

    @staticmethod
    def forward(ctx, q_xyz, c_xyz, idx, feat):
        out, w = K.interp3(q_xyz, c_xyz, idx, feat)
        ctx.save_for_backward(idx, w)
        ctx.s = c_xyz.shape[1]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, w = ctx.saved_tensors
        off, perm = _csr(idx, ctx.s)
        g = K.scatter_rows_csr(grad_out.contiguous(), w, off, perm, ctx.s, 3)
        return None, None, None, g


def interp3(q_xyz, c_xyz, idx, feat) -> torch.Tensor:
    return _Interp3.apply(q_xyz.detach().contiguous(), c_xyz.detach().contiguous(), _as_i32(idx), feat.contiguous())


def interp3_composite(q_xyz, c_xyz, idx, feat) -> torch.Tensor:
    """Same arithmetic from differentiable primitives (used when the candidate coordinates
    require grad, i.e. PointWarping in training: pointconv_util.py:2131-2139)."""
    B, N, _ = q_xyz.shape
    rel = gather_rows(c_xyz, idx) - q_xyz.view(B, N, 1, 3)
    dist = torch.norm(rel, dim=3).clamp(min=1e-10)
    inv = 1.0 / dist
    weight = inv / torch.sum(inv, dim=2, keepdim=True)
    return torch.sum(weight.view(B, N, 3, 1) * gather_rows(feat, idx), dim=2)


# ------------------------------------------------------------------- fused linear layers
_PACK_CACHE = _LRU(512)
_AFFINE_CACHE = _LRU(512)


def _packed_weight_cached(w2d: torch.Tensor, mode: int = 0, d: int = 0, wn: int = 0) -> torch.Tensor:
    """bf16 hi/lo chunk images of a [N,K] weight, cached per (storage, version)."""
    key = (mode, d, wn, _wkey(w2d))
    hit = _PACK_CACHE.get(key)
    if hit is not None:
        return hit[0]
    packed = K.pack_weight(w2d.detach().contiguous(), mode, d, wn)
    _PACK_CACHE.put(key, (packed, w2d))
    return packed


_packed_weight = _packed_weight_cached


def _fold_affine(bias: Optional[torch.Tensor], bn: Optional[torch.nn.Module]):
    """(scale, shift) of the epilogue: Linear bias and eval-mode BatchNorm folded together,
    y = (x W^T + b - mean) * gamma / sqrt(var + eps) + beta."""
    if bn is None:
        return None, (None if bias is None else bias.detach())
    parts = (bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple(None if t is None else _wkey(t) for t in parts) + (bn.eps,)
    hit = _AFFINE_CACHE.get(key)
    if hit is not None:
        return hit[0], hit[1]
    with torch.no_grad():
        scale = torch.rsqrt(bn.running_var + bn.eps)
        if bn.weight is not None:
            scale = scale * bn.weight
        shift = -bn.running_mean * scale
        if bias is not None:
            shift = shift + bias * scale
        if bn.bias is not None:
            shift = shift + bn.bias
        scale, shift = scale.contiguous(), shift.contiguous()
    _AFFINE_CACHE.put(key, (scale, shift, parts))
    return scale, shift


def fused_linear_available(x: torch.Tensor, weight: torch.Tensor, bias, bn) -> bool:
    """The tcgen05 / SIMT fused layer is forward-only: use it when nothing needs a gradient and any
    BatchNorm is in eval mode."""
    if bn is not None and (bn.training or not bn.track_running_stats):
        return False
    if not x.is_cuda or x.dtype != torch.float32:
        return False
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
        return False
    return True


def fused_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 bn: Optional[torch.nn.Module] = None, slope: float = 1.0, clamp=None,
                 residual: Optional[torch.Tensor] = None, cache_weight: bool = True,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[..., N] = clamp(leaky(bn(x[..., K] W^T + b), slope)) + residual in ONE kernel.
    weight [N,K] (nn.Linear) or [N,K,1(,1)] (1x1 conv).
    ``out``: optional row-strided [..., N] view to write into (a column block or a batch half of a wider activation
    buffer, ``ops.row_stride``); ``x`` may be such a view as well - the tensor-core layers take both row strides, so the
    inference forward needs no torch.cat around its 1x1 convolutions."""
    w2d = weight.reshape(weight.shape[0], -1)
    n, k = w2d.shape
    scale, shift = _fold_affine(bias, bn)
    lo, hi = (1.0, 0.0) if clamp is None else (float(clamp[0]), float(clamp[1]))
    res = None if residual is None else residual.contiguous()
    _packed_weight = _packed_weight_cached if cache_weight else (lambda w: K.pack_weight(w.detach().contiguous(), 0, 0, 0))
    tc = not (k < 16 or n < 16 or k % 4 != 0)
    if tc and (out is not None or not x.is_contiguous()):
        ldx = ops.row_stride(x)
        if ldx is None or ldx % 4 != 0 or x.data_ptr() % 16 != 0:
            x = x.contiguous()
        if out is None:
            out = torch.empty(tuple(x.shape[:-1]) + (n,), dtype=torch.float32, device=x.device)
        return ops.linear_tc_into(x, _packed_weight(w2d), n, out, scale, shift, slope, lo, hi, res)
    x = x.contiguous()
    if not tc:
        y = K.linear_simt(x, w2d.detach().contiguous(), scale, shift, slope, lo, hi, res)
    else:
        # any width in ONE launch: layers wider than 256 outputs (level3_1: 256 -> 512; the input gradients of the
        # PointConv linears, up to 8240 columns) run as (row tile, 128-column block) work items against a weight packed
        # in whole column blocks (csrc/linear_tc.cu) - no per-block packing / launching from the host
        y = K.linear_tc(x, _packed_weight(w2d), n, scale, shift, slope, lo, hi, res)
    if out is not None:
        out.copy_(y)
        return out
    return y


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b for the TRAINING path: forward and the input gradient dX = dY W run on tcgen05 (bf16 hi/lo split,
    fp32-level accuracy - the same kernel as inference); dW = dY^T X on tcgen05 with MN-major operand tiles and a
    deterministic split over the rows (csrc/dw_tc.cu); db = column sums stays a torch reduction.  (torch's own fp32
    matmul runs on the CUDA cores: cutlass_80_simt_sgemm was 28 % of the KD training step.)"""

    @staticmethod
    def forward(ctx, x, w2d, bias, slope=1.0):
        # slope != 1: LeakyReLU / ReLU applied in the tcgen05 epilogue (no separate activation kernel, no pre-activation
        # tensor); the backward masks the incoming gradient with the sign of the OUTPUT (same sign as the pre-activation
        # for slope > 0; for ReLU the mask is y > 0 exactly like threshold_backward)
        y = fused_linear(x.detach(), w2d.detach(), None if bias is None else bias.detach(), None, slope)
        ctx.slope = float(slope)
        if ctx.slope != 1.0:
            ctx.save_for_backward(x, w2d, y)
        else:
            ctx.save_for_backward(x, w2d)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w2d = ctx.saved_tensors[0], ctx.saved_tensors[1]
        gy = gy.contiguous()
        if ctx.slope != 1.0:
            gy = torch.ops.aten.leaky_relu_backward(gy, ctx.saved_tensors[2], ctx.slope, True)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = fused_linear(gy, w2d.detach().t().contiguous(), None, cache_weight=False)   # a one-shot tensor: never cached
        g2 = gy.reshape(-1, gy.shape[-1])
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] and USE_TC_DW:
            # tcgen05, MN-major operands; the bias gradient comes out of the same launch as an extra all-ones column of
            # X - unless that column would open a new 256-wide output tile (K % 256 == 0): then db is a torch reduction
            fuse_b = want_b and x.shape[-1] % 256 != 0
            gw, gb = K.linear_dw(g2, x.detach().reshape(-1, x.shape[-1]).contiguous(), fuse_b)
            gb = gb if fuse_b else (g2.sum(0) if want_b else None)
        else:
            if ctx.needs_input_grad[1]:
                gw = g2.t().mm(x.detach().reshape(-1, x.shape[-1]))
            if want_b:
                gb = g2.sum(0)
        return gx, gw, gb, None


class _LinearSmall(torch.autograd.Function):
    """y = x W^T + b for layers with a TINY channel count on one side (3 -> 32 positional encodings over B*N*K rows,
    64 -> 3 flow heads): forward and dX on the CUDA-core kernel (kdpc_linear_simt), dW / db on the tcgen05 weight-gradient
    kernel (zero-padded to its 128 x 16 minimum tile: the reduction over millions of rows is what costs).  torch's
    addmm / mm for these shapes land on SIMT sgemm kernels at ~600 us per call."""

    @staticmethod
    def forward(ctx, x, w2d, bias):
        ctx.save_for_backward(x, w2d)
        ctx.has_bias = bias is not None
        return K.linear_simt(x.detach().contiguous(), w2d.detach().contiguous(), None, None if bias is None else bias.detach(), 1.0, 1.0, 0.0, None)

    @staticmethod
    def backward(ctx, gy):
        x, w2d = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = K.linear_simt(gy, w2d.detach().t().contiguous(), None, None, 1.0, 1.0, 0.0, None)
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or want_b:
            gw, gb = K.linear_dw(gy.reshape(-1, gy.shape[-1]), x.detach().reshape(-1, x.shape[-1]).contiguous(), want_b)
            gb = gb if want_b else None
            gw = gw if ctx.needs_input_grad[1] else None
        return gx, gw, gb


def linear_small_autograd_available(x: torch.Tensor, w2d: torch.Tensor) -> bool:
    n, k = w2d.shape
    return (USE_TC_TRAINING and USE_TC_DW and x.is_cuda and x.dtype == torch.float32 and min(n, k) < 16 and max(n, k) <= 256
            and x.numel() // max(k, 1) >= 4096)


def linear_small_autograd(x: torch.Tensor, w2d: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    return _LinearSmall.apply(x, w2d, bias)


def linear_tc_autograd_available(x: torch.Tensor, w2d: torch.Tensor) -> bool:
    n, k = w2d.shape
    return (USE_TC_TRAINING and x.is_cuda and x.dtype == torch.float32 and k >= 16 and n >= 16 and k % 4 == 0 and n % 4 == 0
            and x.numel() // max(k, 1) >= 128)


def linear_tc_autograd(x: torch.Tensor, w2d: torch.Tensor, bias: Optional[torch.Tensor], slope: float = 1.0) -> torch.Tensor:
    return _LinearTC.apply(x, w2d, bias, slope)


USE_TC_TRAINING = os.environ.get("KDPC_TC_TRAINING", "1") != "0"
USE_TC_DW = os.environ.get("KDPC_TC_DW", "1") != "0"           # A/B: dW = dY^T X on tcgen05 (csrc/dw_tc.cu) or torch.mm


# ------------------------------------------------------------------- fused PointConv / cost volume
_WN_CACHE = _LRU(256)


def _weightnet_host_params(convs):
    """The 248 parameters of a WeightNet(3, 16, hidden=[8, 8]) as a host float list in kernel-launch order
    (w1 b1 w2 b2 w3 b3); one device->host copy per weight version (the fused kernel takes them as
    launch parameters so that every FFMA reads its weight from the constant bank)."""
    ts = [t for c in convs for t in (c.weight, c.bias)]
    key = tuple(_wkey(t) for t in ts)
    hit = _WN_CACHE.get(key)
    if hit is not None:
        return hit[0]
    with torch.no_grad():
        flat = torch.cat([t.detach().reshape(-1).float() for t in ts]).cpu().tolist()
    _WN_CACHE.put(key, (flat, ts))
    return flat


def fused_pointconv_available(weightnet, linear, bn, nsample: int, feats: torch.Tensor) -> bool:
    """Inference-only fused PointConv: K in {9, 16}, WeightNet(3->8->8->16) without BN, D % 4 == 0, Cout <= 256."""
    c = weightnet.mlp_convs
    if weightnet.bn or len(c) != 3 or (c[0].in_channels, c[0].out_channels, c[1].out_channels, c[2].out_channels) != (3, 8, 8, 16):
        return False
    if nsample not in FUSED_POINTCONV_K or feats is None or feats.shape[2] % 4 != 0 or linear.out_features > 256:
        return False
    if linear.in_features != 16 * (feats.shape[2] + 3):
        return False
    if torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in weightnet.parameters())):
        return False
    return fused_linear_available(feats, linear.weight, linear.bias, bn)


FUSED_POINTCONV_K = (9, 16)


def morton_order(xyz_d: torch.Tensor):
    """int32 [B,N] view (row stride = one cloud's block) of the Morton order of a cloud that has ALREADY been
    spatially sorted for a kNN call (cached), else None.  No kernel is launched here."""
    if not _CACHE_ENABLED or not xyz_d.is_contiguous():
        return None
    hit = _SORT_CACHE.get(_tkey(xyz_d))
    if hit is None:
        return None
    B, N, _ = xyz_d.shape
    L = _lib.lib()
    stride, off = L.kdpc_spatial_sort_order_stride(N), L.kdpc_spatial_sort_order_offset(N)
    buf = hit[0]
    if buf.numel() != B * stride * 4:
        return None
    return buf.view(torch.int32).view(B, stride)[:, off:off + N]


def fused_pointconv(cand_xyz, query_xyz, feats, idx, weightnet, linear, bn, slope: float) -> torch.Tensor:
    """[B,S,Cout] = act(bn(Linear(sum_k [feats[idx], rel_xyz] (x) WeightNet(rel_xyz)))) in ONE kernel.
    The queries are processed in Morton order when the kNN that produced ``idx`` left one behind (same results;
    spatially coherent tiles share their neighbours, so the gathers hit in L1)."""
    d = feats.shape[2]
    wp = _packed_weight(linear.weight, 1, d, 16)
    scale, shift = _fold_affine(linear.bias, bn)
    query_xyz = query_xyz.contiguous()
    return K.pointconv_fused(cand_xyz.contiguous(), query_xyz, feats.contiguous(), _as_i32(idx),
                             _weightnet_host_params(weightnet.mlp_convs), wp, linear.out_features, scale, shift, slope,
                             morton_order(query_xyz) if USE_MORTON_ORDER else None)


USE_MORTON_ORDER = os.environ.get("KDPC_PC_ORDER", "1") != "0"


def fused_costvol(xyz1, xyz2, p1, p2, idx, pos, slope_pre: float, conv, slope_post: float) -> torch.Tensor:
    """max_k act(conv(act(p2[idx] + p1 + pos(xyz2[idx] - xyz1)))) -> [B,N1,Dout] in ONE kernel."""
    d = p1.shape[2]
    w2d = conv.weight.reshape(conv.weight.shape[0], -1)
    return K.costvol_fused(xyz1.contiguous(), xyz2.contiguous(), p1.contiguous(), p2.contiguous(), _as_i32(idx),
                           pos.weight.detach().reshape(d, 3).contiguous(), pos.bias.detach(), slope_pre,
                           _packed_weight(w2d), w2d.shape[0], None if conv.bias is None else conv.bias.detach(),
                           slope_post)


class _CostVolFn(torch.autograd.Function):
    """Fused cost-volume half in its folded form, out = act2(max_k(W act1(p2q[idx] + p1q) + b)) (K = 32, D = D' = 32 or 64):
    forward = the tcgen05 inference kernel (costvol_tc.cu), backward = ONE recomputing kernel in which only the arg-max
    neighbour of every (point, channel) receives a gradient (costvol_grad.cu) + the deterministic CSR scatter of the
    gathered rows' gradients.  Nothing of size [B,N,K,*] is saved."""

    @staticmethod
    def forward(ctx, p1q, p2q, idx, w2d, bias, slope_pre, slope_post):
        ctx.save_for_backward(p1q, p2q, idx, w2d, bias)
        ctx.slopes = (float(slope_pre), float(slope_post))
        d = p1q.shape[2]
        zero_w = p1q.new_zeros((d, 3))
        zero_b = p1q.new_zeros((d,))
        xyz1 = p1q.new_zeros((p1q.shape[0], p1q.shape[1], 3))
        xyz2 = p1q.new_zeros((p2q.shape[0], p2q.shape[1], 3))
        return K.costvol_fused(xyz1, xyz2, p1q.detach(), p2q.detach(), idx, zero_w, zero_b, ctx.slopes[0],
                               K.pack_weight(w2d.detach().contiguous(), 0, 0, 0), w2d.shape[0],
                               None if bias is None else bias.detach(), ctx.slopes[1])

    @staticmethod
    def backward(ctx, g):
        p1q, p2q, idx, w2d, bias = ctx.saved_tensors
        g1, grows, gw, gb = K.costvol_grad(p1q, p2q, idx, w2d.detach().contiguous(), None if bias is None else bias.detach(),
                                           ctx.slopes[0], ctx.slopes[1], g.contiguous())
        g2 = None
        if ctx.needs_input_grad[1]:
            off, perm = _csr(idx, p2q.shape[1])
            g2 = K.scatter_rows_csr(grows, None, off, perm, p2q.shape[1], 1)
        return g1, g2, None, gw, (gb if bias is not None else None), None, None


USE_FUSED_COSTVOL_GRAD = os.environ.get("KDPC_COSTVOL_GRAD", "1") != "0"      # A/B: 0 = autograd over the unfused op chain


def costvol_autograd_available(points1: torch.Tensor, idx: torch.Tensor, conv, slope_pre: float) -> bool:
    w = conv.weight
    d = points1.shape[2]
    return (USE_TC_TRAINING and USE_FUSED_COSTVOL_GRAD and points1.is_cuda and points1.dtype == torch.float32
            and d in (32, 64) and idx.shape[2] == 32 and w.shape[0] == d and w.reshape(d, -1).shape[1] == d
            and 0.0 <= slope_pre < 1.0)


def costvol_autograd(xyz1, xyz2, points1, points2, idx, pos, slope_pre: float, conv, slope_post: float) -> torch.Tensor:
    """Differentiable max_k act(conv(act(p2[idx] + p1 + pos(xyz2[idx] - xyz1)))) for the training path.  The positional
    layer is linear, so it is folded into the point features once per POINT by small differentiable ops
    (p2q = p2 + pos_w xyz2, p1q = p1 + pos_b - pos_w xyz1: autograd un-folds their gradients) and the [B,N,K,*] part runs
    as _CostVolFn."""
    d = points1.shape[2]
    pw = pos.weight.reshape(d, 3)
    lin = linear_small_autograd if linear_small_autograd_available(xyz1, pw) else (lambda x, w, b: torch.nn.functional.linear(x, w, b))
    p2q = points2 + lin(xyz2.contiguous(), pw, None)
    p1q = points1 + pos.bias - lin(xyz1.contiguous(), pw, None)
    w2d = conv.weight.reshape(conv.weight.shape[0], -1)
    return _CostVolFn.apply(p1q.contiguous(), p2q.contiguous(), _as_i32(idx), w2d, conv.bias, slope_pre, slope_post)


# ------------------------------------------------------------------- WeightNet (training path)
class _WeightNetFn(torch.autograd.Function):
    """relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3) on the first 3 columns of ``rel`` (pointconv_util.py:204-215, bn=False):
    forward = the fused inference kernel, backward = ONE kernel that recomputes the forward per row and reduces the six
    parameter gradients deterministically (csrc/weightnet_grad.cu).  Only ``rel`` is saved."""

    @staticmethod
    def forward(ctx, rel, w1, b1, w2, b2, w3, b3):
        ctx.save_for_backward(rel, w1, b1, w2, b2, w3, b3)
        return K.weightnet(rel, w1, b1, w2, b2, w3, b3)

    @staticmethod
    def backward(ctx, g):
        rel, w1, b1, w2, b2, w3, b3 = ctx.saved_tensors
        want_x = ctx.needs_input_grad[0]
        gw1, gb1, gw2, gb2, gw3, gb3, gx3 = K.weightnet_grad(rel, g.contiguous(), w1, b1, w2, b2, w3, b3, want_x)
        g_rel = None
        if want_x:
            g_rel = torch.zeros_like(rel)
            g_rel[..., :3] = gx3
        return g_rel, gw1, gb1, gw2, gb2, gw3, gb3


def weightnet_autograd(rel: torch.Tensor, convs) -> torch.Tensor:
    """Differentiable WeightNet(3 -> 8 -> 8 -> W in {8, 16}) on a point-major [..., C>=3] tensor."""
    return _WeightNetFn.apply(rel.contiguous(), convs[0].weight.reshape(8, 3), convs[0].bias, convs[1].weight.reshape(8, 8),
                              convs[1].bias, convs[2].weight.reshape(-1, 8), convs[2].bias)


USE_FUSED_WEIGHTNET_GRAD = os.environ.get("KDPC_WN_GRAD", "1") != "0"


# ------------------------------------------------------------------- PointConv pieces
class _PointConvAgg(torch.autograd.Function):
    """out[b,s,c*W+w] = sum_k grouped[b,s,k,c] * wn[b,s,k,w]   (pointconv_util.py:249)."""

    @staticmethod
    def forward(ctx, grouped, wn):
        ctx.save_for_backward(grouped, wn)
        return K.pointconv_agg(grouped, wn)

    @staticmethod
    def backward(ctx, grad_out):
        grouped, wn = ctx.saved_tensors
        B, S, Kn, C = grouped.shape
        W = wn.shape[3]
        want_g, want_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if W == 16 and grad_out.is_cuda:                       # one kernel for both gradients (csrc/pointconv.cu)
            gg, gw = K.pointconv_agg_grad(grouped, wn, grad_out.contiguous(), want_g, want_w)
            return (gg if want_g else None), (gw if want_w else None)
        g = grad_out.reshape(B * S, C, W)
        g_grouped = g_wn = None
        if want_g:
            g_grouped = torch.bmm(wn.reshape(B * S, Kn, W), g.transpose(1, 2)).view(B, S, Kn, C)
        if want_w:
            g_wn = torch.bmm(grouped.reshape(B * S, Kn, C), g).view(B, S, Kn, W)
        return g_grouped, g_wn


def pointconv_agg(grouped: torch.Tensor, wn: torch.Tensor) -> torch.Tensor:
    return _PointConvAgg.apply(grouped.contiguous(), wn.contiguous())


# ------------------------------------------------------------------- fused losses (a18, a19)
class _KDLoss(torch.autograd.Function):
    """loss = sum_t w_t * multiScaleLoss(preds, target_t) + sum_h 0.5 * hw_h * |fs_h - ft_h|^2; the kernels write
    the gradients in the same pass, backward only scales them by the incoming gradient."""

    @staticmethod
    def forward(ctx, meta, *tensors):
        ns, nh, fps_idxs, targets, alpha, weights, hint_ft, hint_w, point_major = meta
        preds, hint_fs = list(tensors[:ns]), list(tensors[ns:ns + nh])
        want = any(ctx.needs_input_grad[1:])
        loss, g_pred, g_hint = K.kd_loss(preds, list(fps_idxs), list(targets), list(alpha), list(weights), point_major,
                                         hint_fs, list(hint_ft), list(hint_w), want)
        ctx.save_for_backward(*g_pred, *g_hint)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        return (None,) + tuple(g * grad_out for g in ctx.saved_tensors)


def kd_loss(pred_flows, fps_idxs, targets, weights, alpha, hints=()) -> torch.Tensor:
    """pred_flows: [B,3,N_i] (the model's outputs; permuted views of point-major tensors are used in place);
    targets: [B,N0,3] tensors (never differentiated: ground truth, or the teacher's flow computed under
    no_grad); weights: one float per target, INCLUDING the 1/B of the batch mean; hints: (student_feat,
    teacher_feat, weight) triples."""
    preds = list(pred_flows)
    point_major = all(p.dim() == 3 and p.stride(1) == 1 and p.permute(0, 2, 1).is_contiguous() for p in preds)
    preds = [p.permute(0, 2, 1) for p in preds] if point_major else [p.contiguous() for p in preds]
    fs = [h[0].contiguous() for h in hints]
    ft = [h[1].detach().contiguous() for h in hints]
    for a, b in zip(fs, ft):
        if a.shape != b.shape:                     # what `fs - ft` raises in the reference (loss_functions.py:213-214)
            raise RuntimeError(f"The size of tensor a {tuple(a.shape)} must match the size of tensor b {tuple(b.shape)}")
    meta = (len(preds), len(fs), tuple(_as_i32(i) for i in fps_idxs), tuple(t.detach().contiguous() for t in targets),
            tuple(float(a) for a in alpha), tuple(float(w) for w in weights), tuple(ft), tuple(float(h[2]) for h in hints),
            point_major)
    return _KDLoss.apply(meta, *preds, *fs)
