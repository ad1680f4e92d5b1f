"""ORACLE (test infrastructure, never on the product path): torch-CPU restatement of the reference's self-supervised
loss family, models_bid_pointconv.py:565-677 (curvature, computeChamfer, curvatureWarp, computeSmooth,
interpolateCurvature, multiScaleChamferSmoothCurvature).  Same op chain: the full [B,N,M] matrix of
``square_distance`` (matmul expansion) and ``torch.topk`` on it.  Pinned bit-for-bit against the unmodified reference
functions by tests/make_golden_selfsup.py -> tests/golden/selfsup.npz.
"""
import torch

from .layers_ref import index_points_group, square_distance


def _pm(x):
    return x.permute(0, 2, 1)


def _nearest(query_pm, cand_pm, k):
    """the reference's neighbour search: full square_distance matrix + unsorted topk -> (dist, idx) [B,S,k]"""
    return torch.topk(square_distance(query_pm, cand_pm), k, dim=-1, largest=False, sorted=False)


def _laplacian(neigh_of_pm, values_pm):
    """sum over the 10 nearest neighbours (found in ``neigh_of_pm``) of (value_j - value_i), / 9  (:565-572, :592-599)"""
    _, kidx = _nearest(neigh_of_pm, neigh_of_pm, 10)
    return torch.sum(index_points_group(values_pm, kidx) - values_pm.unsqueeze(2), dim=2) / 9.0


def curvature(pc):
    """:565-572.  pc [B,3,N] -> [B,N,3]"""
    return _laplacian(_pm(pc), _pm(pc))


def curvature_warp(pc, warped_pc):
    """:592-599: neighbourhoods of pc, coordinates of warped_pc"""
    return _laplacian(_pm(pc), _pm(warped_pc))


def compute_chamfer(pc1, pc2):
    """:574-590.  pc1 [B,3,N], pc2 [B,3,M] -> dist1 [B,N], dist2 [B,M]: row / column minima of ONE distance matrix"""
    sqrdist12 = square_distance(_pm(pc1), _pm(pc2))
    dist1, _ = torch.topk(sqrdist12, 1, dim=-1, largest=False, sorted=False)
    dist2, _ = torch.topk(sqrdist12, 1, dim=1, largest=False, sorted=False)
    return dist1.squeeze(2), dist2.squeeze(1)


def compute_smooth(pc1, pred_flow):
    """:601-616 -> [B,N]: mean distance of a point's flow to the flows of its 9 nearest neighbours (self included), / 8"""
    flow = _pm(pred_flow)
    _, kidx = _nearest(_pm(pc1), _pm(pc1), 9)
    return torch.norm(index_points_group(flow, kidx) - flow.unsqueeze(2), dim=3).sum(dim=2) / 8.0


def interpolate_curvature(pc1, pc2, pc2_curvature):
    """:618-638.  pc2_curvature [B,M,3] (point-major, as `curvature` returns it): inverse squared-distance weights over 5"""
    B, _, N = pc1.shape
    dist, knn_idx = _nearest(_pm(pc1), _pm(pc2), 5)
    grouped = index_points_group(pc2_curvature, knn_idx)
    norm = torch.sum(1.0 / (dist + 1e-8), dim=2, keepdim=True)
    weight = (1.0 / (dist + 1e-8)) / norm
    return torch.sum(weight.view(B, N, 5, 1) * grouped, dim=2)


def multi_scale_chamfer_smooth_curvature(pc1, pc2, pred_flows, alpha=(0.02, 0.04, 0.08, 0.16)):
    """:640-677.  Returns (total, chamfer, curvature, smoothness), each shape [1]."""
    f_curvature, f_smoothness, f_chamfer = 0.3, 1.0, 1.0
    chamfer_loss, smoothness_loss, curvature_loss = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    for i in range(len(pred_flows)):
        cur_pc1, cur_pc2, cur_flow = pc1[i], pc2[i], pred_flows[i]
        cur_pc2_curvature = curvature(cur_pc2)
        cur_pc1_warp = cur_pc1 + cur_flow
        dist1, dist2 = compute_chamfer(cur_pc1_warp, cur_pc2)
        moved_pc1_curvature = curvature_warp(cur_pc1, cur_pc1_warp)
        chamfer = dist1.sum(dim=1).mean() + dist2.sum(dim=1).mean()
        smooth = compute_smooth(cur_pc1, cur_flow).sum(dim=1).mean()
        inter = interpolate_curvature(cur_pc1_warp, cur_pc2, cur_pc2_curvature)
        curv = torch.sum((inter - moved_pc1_curvature) ** 2, dim=2).sum(dim=1).mean()
        chamfer_loss = chamfer_loss + alpha[i] * chamfer
        smoothness_loss = smoothness_loss + alpha[i] * smooth
        curvature_loss = curvature_loss + alpha[i] * curv
    total = f_chamfer * chamfer_loss + f_curvature * curvature_loss + f_smoothness * smoothness_loss
    return total, chamfer_loss, curvature_loss, smoothness_loss
