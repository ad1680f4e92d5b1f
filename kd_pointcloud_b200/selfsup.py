"""The reference's self-supervised loss family (models_bid_pointconv.py:565-677: chamfer, flow smoothness, curvature)
on the kdpc kernels.

The reference builds the full [B,N,M] matrix of ``square_distance`` and runs ``torch.topk`` on it for each of the six
neighbour searches per scale (K = 1, 1, 5, 9, 10, 10).  Here every search is the exact tile-pruned kNN kernel
(``kdpc_knn``: same distances, ascending (distance, index)), the neighbour rows come from ``gather_rows`` (deterministic
CSR scatter-add backward) and only the few elementwise steps of the formulas are torch ops, so the whole family is
differentiable with respect to the predicted flow exactly as in the reference.

Distances that enter the loss by VALUE (chamfer, interpolation weights) are taken from the kernel - they carry the
reference's own rounding of the matmul expansion - while their gradient is that of |p - q|^2 (what autograd derives
for the reference's expression): ``_knn_sqdist``.

Names, argument layouts ([B,3,N]) and return values follow the reference functions.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import functional as KF
from . import ops  # noqa: F401  (registers torch.ops.kdpc)

K = torch.ops.kdpc


def _pm(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 1).contiguous()


def _knn_sqdist(query: torch.Tensor, cand: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """idx int32 [B,S,k] and squared distances [B,S,k] of the k nearest candidates of every query (point-major
    [B,*,3] inputs).  Forward value: the kernel's (reference rounding); gradient: d|q - c|^2 to both clouds."""
    idx, dist = K.knn_dist(query.detach().contiguous(), cand.detach().contiguous(), k)
    if not (torch.is_grad_enabled() and (query.requires_grad or cand.requires_grad)):
        return idx, dist
    diff = query.unsqueeze(2) - KF.gather_rows(cand, idx)                 # [B,S,k,3], differentiable
    soft = (diff * diff).sum(-1)
    return idx, soft + (dist - soft).detach()


def curvature(pc: torch.Tensor) -> torch.Tensor:
    """models_bid_pointconv.py:565-572.  pc [B,3,N] -> [B,N,3]"""
    p = _pm(pc)
    idx = KF.knn_idx(10, p, p)
    return (KF.gather_rows(p, idx) - p.unsqueeze(2)).sum(dim=2) / 9.0


def computeChamfer(pc1: torch.Tensor, pc2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """:574-590.  pc1 [B,3,N], pc2 [B,3,M] -> dist1 [B,N] (nearest pc2 point of every pc1 point), dist2 [B,M]."""
    p1, p2 = _pm(pc1), _pm(pc2)
    _, d1 = _knn_sqdist(p1, p2, 1)
    _, d2 = _knn_sqdist(p2, p1, 1)
    return d1.squeeze(2), d2.squeeze(2)


def curvatureWarp(pc: torch.Tensor, warped_pc: torch.Tensor) -> torch.Tensor:
    """:592-599: neighbourhoods of ``pc``, coordinates of ``warped_pc``."""
    p, w = _pm(pc), _pm(warped_pc)
    idx = KF.knn_idx(10, p, p)
    return (KF.gather_rows(w, idx) - w.unsqueeze(2)).sum(dim=2) / 9.0


def computeSmooth(pc1: torch.Tensor, pred_flow: torch.Tensor) -> torch.Tensor:
    """:601-616 -> [B,N]"""
    p, f = _pm(pc1), _pm(pred_flow)
    idx = KF.knn_idx(9, p, p)
    return torch.norm(KF.gather_rows(f, idx) - f.unsqueeze(2), dim=3).sum(dim=2) / 8.0


def interpolateCurvature(pc1: torch.Tensor, pc2: torch.Tensor, pc2_curvature: torch.Tensor) -> torch.Tensor:
    """:618-638.  pc2_curvature [B,M,3] -> [B,N,3]: inverse-(squared-)distance weights over the 5 nearest pc2 points."""
    p1, p2 = _pm(pc1), _pm(pc2)
    idx, dist = _knn_sqdist(p1, p2, 5)
    inv = 1.0 / (dist + 1e-8)
    weight = inv / inv.sum(dim=2, keepdim=True)
    return (weight.unsqueeze(-1) * KF.gather_rows(pc2_curvature.contiguous(), idx)).sum(dim=2)


SCALE_WEIGHTS = (0.02, 0.04, 0.08, 0.16)       # alpha per pyramid level, finest first (:649)
TERM_WEIGHTS = {"chamfer": 1.0, "curvature": 0.3, "smoothness": 1.0}     # f_chamfer, f_curvature, f_smoothness (:641-643)


def _scale_terms(p1: torch.Tensor, p2: torch.Tensor, flow: torch.Tensor):
    """chamfer, curvature and smoothness terms of ONE pyramid level (:657-673), each a 0-dim tensor (batch mean of
    per-cloud sums)."""
    warped = p1 + flow
    d12, d21 = computeChamfer(warped, p2)
    chamfer = d12.sum(dim=1).mean() + d21.sum(dim=1).mean()
    smooth = computeSmooth(p1, flow).sum(dim=1).mean()
    target = interpolateCurvature(warped, p2, curvature(p2))
    curv = ((target - curvatureWarp(p1, warped)) ** 2).sum(dim=2).sum(dim=1).mean()
    return chamfer, curv, smooth


def multiScaleChamferSmoothCurvature(pc1: Sequence[torch.Tensor], pc2: Sequence[torch.Tensor],
                                     pred_flows: Sequence[torch.Tensor]):
    """:640-677.  Lists of [B,3,N_i] tensors (finest first) -> (total, chamfer, curvature, smoothness), each of shape [1]."""
    dev = pred_flows[0].device
    sums = [torch.zeros(1, device=dev) for _ in range(3)]
    for level, flow in enumerate(pred_flows):
        for acc_i, term in enumerate(_scale_terms(pc1[level], pc2[level], flow)):
            sums[acc_i] = sums[acc_i] + SCALE_WEIGHTS[level] * term
    chamfer_loss, curvature_loss, smoothness_loss = sums
    total = (TERM_WEIGHTS["chamfer"] * chamfer_loss + TERM_WEIGHTS["curvature"] * curvature_loss
             + TERM_WEIGHTS["smoothness"] * smoothness_loss)
    return total, chamfer_loss, curvature_loss, smoothness_loss
