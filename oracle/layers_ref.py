"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle", kind = port) of the reference's torch
layer library for the hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Every function follows the op chain of one reference definition (file:line in its docstring) in
plain functional torch on CPU tensors, with parameters taken from a ``state_dict`` by prefix, so
that it can be compared against (a) the real reference modules imported from /root/reference in
this container (tests/make_golden.py pins it and writes tests/golden/), and (b) the CUDA path.

Pinned: yes — tests/test_oracle_golden.py checks this file against the golden vectors produced by
the unmodified reference code, and tests/make_golden.py asserts equality when it generates them.
The pointnet2 ops the reference only has as CUDA kernels (FPS, gather, group) come from
oracle/kdpc_oracle.c (pinned against oracle/_ref on the GPU box).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LEAKY_RATE = 0.1
_lib = None


def clib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
        _lib = ctypes.CDLL(path)
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np32(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


def _npi(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.int32)


# ------------------------------------------------------------------ pointnet2 ops (C oracle)
def furthest_point_sample(xyz: torch.Tensor, npoint: int, closed_form: bool = False) -> torch.Tensor:
    """pointnet2_utils.py:10-36 -> sampling_gpu.cu:93-209."""
    x = _np32(xyz)
    B, N, _ = x.shape
    idx = np.zeros((B, npoint), dtype=np.int32)
    temp = np.full((B, N), 1e10, dtype=np.float32)
    fn = clib().oracle_fps_closed_form if closed_form else clib().oracle_fps
    fn(B, N, npoint, _fp(x), _fp(temp), _fp(idx))
    return torch.from_numpy(idx)


def three_nn(unknown: torch.Tensor, known: torch.Tensor):
    """pointnet2_utils.py:76-105: returns (sqrt(d2), idx)."""
    u, k = _np32(unknown), _np32(known)
    B, n, _ = u.shape
    m = k.shape[1]
    d2 = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    clib().oracle_three_nn(B, n, m, _fp(u), _fp(k), _fp(d2), _fp(idx))
    return torch.sqrt(torch.from_numpy(d2)), torch.from_numpy(idx)


def three_interpolate(features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    f, i, w = _np32(features), _npi(idx), _np32(weight)
    B, c, m = f.shape
    n = i.shape[1]
    out = np.zeros((B, c, n), dtype=np.float32)
    clib().oracle_three_interpolate(B, c, m, n, _fp(f), _fp(i), _fp(w), _fp(out))
    return torch.from_numpy(out)


def ball_query(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    x, q = _np32(xyz), _np32(new_xyz)
    B, n, _ = x.shape
    m = q.shape[1]
    idx = np.zeros((B, m, nsample), dtype=np.int32)
    clib().oracle_ball_query(B, n, m, ctypes.c_float(radius), nsample, _fp(q), _fp(x), _fp(idx))
    return torch.from_numpy(idx)


def gather_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[B,C,N],[B,M] -> [B,C,M]  (sampling_gpu.cu:8-24)."""
    B, C, _ = features.shape
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(B, C, idx.shape[1]))


def grouping_operation(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[B,C,N],[B,S,K] -> [B,C,S,K]  (group_points_gpu.cu:47-66)."""
    B, C, _ = features.shape
    _, S, Kn = idx.shape
    flat = idx.long().reshape(B, 1, S * Kn).expand(B, C, S * Kn)
    return torch.gather(features, 2, flat).view(B, C, S, Kn)


# ------------------------------------------------------------------ pointconv_util functions
def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """pointconv_util.py:73-94, same three steps."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def square_distance_c(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """Scalar-C evaluation of the same expression (machine independent rounding order)."""
    s, d = _np32(src), _np32(dst)
    B, S, _ = s.shape
    N = d.shape[1]
    out = np.zeros((B, S, N), dtype=np.float32)
    clib().oracle_square_distance(B, S, N, _fp(s), _fp(d), _fp(out))
    return torch.from_numpy(out)


def knn_point(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, impl: str = "c") -> torch.Tensor:
    """pointconv_util.py:96-107.  impl='torch' is the reference op chain (matmul expansion +
    topk, order unspecified); impl='c' selects by (distance, index) on the same distances and
    returns them ascending (what the CUDA kernel promises)."""
    if impl == "torch":
        sq = square_distance(new_xyz, xyz)
        return torch.topk(sq, nsample, dim=-1, largest=False, sorted=False)[1]
    q, c = _np32(new_xyz), _np32(xyz)
    B, S, _ = q.shape
    N = c.shape[1]
    idx = np.zeros((B, S, nsample), dtype=np.int32)
    clib().oracle_knn(B, S, N, nsample, _fp(q), _fp(c), _fp(idx), None)
    return torch.from_numpy(idx).long()


def knn_with_dist(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor):
    q, c = _np32(new_xyz), _np32(xyz)
    B, S, _ = q.shape
    N = c.shape[1]
    idx = np.zeros((B, S, nsample), dtype=np.int32)
    dist = np.zeros((B, S, nsample), dtype=np.float32)
    clib().oracle_knn(B, S, N, nsample, _fp(q), _fp(c), _fp(idx), _fp(dist))
    return torch.from_numpy(idx), torch.from_numpy(dist)


def index_points_gather(points: torch.Tensor, fps_idx: torch.Tensor) -> torch.Tensor:
    """pointconv_util.py:109-120: [B,N,C],[B,S] -> [B,S,C]."""
    return gather_operation(points.permute(0, 2, 1).contiguous(), fps_idx).permute(0, 2, 1).contiguous()


def index_points_group(points: torch.Tensor, knn_idx: torch.Tensor) -> torch.Tensor:
    """pointconv_util.py:122-133: [B,N,C],[B,S,K] -> [B,S,K,C]."""
    return grouping_operation(points.permute(0, 2, 1).contiguous(), knn_idx.int()).permute(0, 2, 3, 1)


def group_query(nsample, s_xyz, xyz, s_points, knn_impl="c"):
    """pointconv_util.py:159-182 (group() at :135-157 is the s_xyz == xyz case)."""
    B, S, C = xyz.shape
    idx = knn_point(nsample, s_xyz, xyz, knn_impl)
    rel = index_points_group(s_xyz, idx) - xyz.view(B, S, 1, C)
    if s_points is None:
        return rel, rel
    return torch.cat([rel, index_points_group(s_points, idx)], dim=-1), rel


def _leaky(x):
    return F.leaky_relu(x, LEAKY_RATE)


def conv1d(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, act: bool = True) -> torch.Tensor:
    """Conv1d wrapper (pointconv_util.py:20-36) when act, bare nn.Conv1d otherwise; x [B,Cin,N]."""
    key = prefix + (".composed_module.0" if act else "")
    y = F.conv1d(x, sd[key + ".weight"], sd[key + ".bias"])
    return _leaky(y) if act else y


def conv2d(sd, prefix, x, act=True):
    key = prefix + (".composed_module.0" if act else "")
    y = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"])
    return _leaky(y) if act else y


def weightnet(sd, prefix, localized_xyz: torch.Tensor) -> torch.Tensor:
    """WeightNet.forward, pointconv_util.py:205-215 (bn=False): ReLU after each 1x1 conv. [B,3,K,N]."""
    w = localized_xyz
    i = 0
    while f"{prefix}.mlp_convs.{i}.weight" in sd:
        w = F.relu(F.conv2d(w, sd[f"{prefix}.mlp_convs.{i}.weight"], sd[f"{prefix}.mlp_convs.{i}.bias"]))
        i += 1
    return w


def _pointconv_tail(sd, prefix, new_points, rel, B, S, bn: bool, training: bool):
    """Shared tail of PointConv / PointConvD (pointconv_util.py:246-256, :434-446)."""
    weights = weightnet(sd, prefix + ".weightnet", rel.permute(0, 3, 2, 1))
    x = torch.matmul(new_points.permute(0, 1, 3, 2), weights.permute(0, 3, 2, 1)).reshape(B, S, -1)
    x = F.linear(x, sd[prefix + ".linear.weight"], sd[prefix + ".linear.bias"])
    x = x.permute(0, 2, 1)
    if bn:
        x = F.batch_norm(x, sd[prefix + ".bn_linear.running_mean"].clone(), sd[prefix + ".bn_linear.running_var"].clone(),
                         sd[prefix + ".bn_linear.weight"], sd[prefix + ".bn_linear.bias"], training, 0.1, 1e-5)
    return _leaky(x)


def pointconv(sd, prefix, nsample, xyz, points, bn=False, training=False, knn_impl="c"):
    """PointConv.forward, pointconv_util.py:231-258. xyz [B,3,N], points [B,D,N] -> [B,Cout,N]."""
    B, _, N = xyz.shape
    xyz_t, pts_t = xyz.permute(0, 2, 1), points.permute(0, 2, 1)
    new_points, rel = group_query(nsample, xyz_t, xyz_t, pts_t, knn_impl)
    return _pointconv_tail(sd, prefix, new_points, rel, B, N, bn, training)


def pointconvd(sd, prefix, npoint, nsample, xyz, points, knn_impl="c"):
    """PointConvD.forward, pointconv_util.py:414-446 -> (new_xyz [B,3,S], feats [B,Cout,S], fps_idx)."""
    B = xyz.shape[0]
    xyz_t, pts_t = xyz.permute(0, 2, 1).contiguous(), points.permute(0, 2, 1)
    fps_idx = furthest_point_sample(xyz_t, npoint)
    new_xyz = index_points_gather(xyz_t, fps_idx)
    new_points, rel = group_query(nsample, xyz_t, new_xyz, pts_t, knn_impl)
    return new_xyz.permute(0, 2, 1), _pointconv_tail(sd, prefix, new_points, rel, B, npoint, False, False), fps_idx


def knn_point_feat(nsample: int, feats: torch.Tensor, new_feats: torch.Tensor) -> torch.Tensor:
    """knn_point on C-dimensional FEATURES (CrossLayerLightFG, pointconv_util.py:1905): the reference's matmul
    expansion, selected by ascending (distance, index) (stable sort) instead of topk's unspecified order."""
    sq = square_distance(new_feats, feats)
    return torch.sort(sq, dim=-1, stable=True)[1][..., :nsample]


def cross(sd, nsample, xyz1, xyz2, points1, points2, pos_prefix, mlp_prefix, knn_impl="c", idx=None):
    """CrossLayerLight.cross, pointconv_util.py:1826-1850 (bn = Identity).  ``idx``: a neighbourhood built by the caller."""
    B, C, N1 = xyz1.shape
    D1 = points1.shape[1]
    x1, x2 = xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1)
    p1, p2 = points1.permute(0, 2, 1), points2.permute(0, 2, 1)
    if idx is None:
        idx = knn_point(nsample, x2, x1, knn_impl)
    nsample = idx.shape[2]
    direction = index_points_group(x2, idx) - x1.view(B, N1, 1, C)
    g2 = index_points_group(p2, idx).permute(0, 3, 2, 1)
    g1 = p1.view(B, N1, 1, D1).repeat(1, 1, nsample, 1).permute(0, 3, 2, 1)
    direction = F.conv2d(direction.permute(0, 3, 2, 1), sd[pos_prefix + ".weight"], sd[pos_prefix + ".bias"])
    x = _leaky(g2 + g1 + direction)
    i = 0
    while f"{mlp_prefix}.{i}.composed_module.0.weight" in sd:
        x = conv2d(sd, f"{mlp_prefix}.{i}", x)
        i += 1
    return F.max_pool2d(x, (x.size(2), 1)).squeeze(2)


def cross_layer_light(sd, prefix, nsample, pc1, pc2, feat1, feat2, knn_impl="c"):
    """CrossLayerLight.forward, pointconv_util.py:1852-1868."""
    t11 = lambda x: conv1d(sd, prefix + ".cross_t11", x, act=False)
    t22 = lambda x: conv1d(sd, prefix + ".cross_t22", x, act=False)
    f1 = cross(sd, nsample, pc1, pc2, t11(feat1), t22(feat2), prefix + ".pos1", prefix + ".mlp1", knn_impl)
    f2 = cross(sd, nsample, pc2, pc1, t11(feat2), t22(feat1), prefix + ".pos1", prefix + ".mlp1", knn_impl)
    f1 = conv1d(sd, prefix + ".cross_t1", f1, act=False)
    f2 = conv1d(sd, prefix + ".cross_t2", f2, act=False)
    f3 = cross(sd, nsample, pc1, pc2, f1, f2, prefix + ".pos2", prefix + ".mlp2", knn_impl)
    return f1, f2, f3


def no_cross_layer_light(sd, prefix, nsample, pc1, pc2, feat1, feat2, knn_impl="c"):
    """NoCrossLayerLight.forward, pointconv_util.py:1276-1331 (bn = Identity)."""
    t1 = conv1d(sd, prefix + ".cross_t1", feat1, act=False)
    t2 = conv1d(sd, prefix + ".cross_t2", feat2, act=False)
    return cross(sd, nsample, pc1, pc2, t1, t2, prefix + ".pos", prefix + ".mlp", knn_impl)


def cross_layer_light_fg(sd, prefix, nsample, pc1, pc2, feat1, feat2, knn1, knn2, knn_impl="c"):
    """CrossLayerLightFG.forward, pointconv_util.py:1871-1957: neighbourhood = nsample/2 nearest in the feature space
    knn1/knn2 followed by nsample/2 nearest in xyz."""
    half = nsample // 2

    def fg(xa, xb, pa, pb, ka, kb, pos, mlp):
        idx = torch.cat([knn_point_feat(half, kb.permute(0, 2, 1), ka.permute(0, 2, 1)),
                         knn_point(half, xb.permute(0, 2, 1), xa.permute(0, 2, 1), knn_impl)], dim=-1)
        return cross(sd, nsample, xa, xb, pa, pb, prefix + pos, prefix + mlp, knn_impl, idx=idx)

    t11 = lambda x: conv1d(sd, prefix + ".cross_t11", x, act=False)
    t22 = lambda x: conv1d(sd, prefix + ".cross_t22", x, act=False)
    f1 = conv1d(sd, prefix + ".cross_t1", fg(pc1, pc2, t11(feat1), t22(feat2), knn1, knn2, ".pos1", ".mlp1"), act=False)
    f2 = conv1d(sd, prefix + ".cross_t2", fg(pc2, pc1, t11(feat2), t22(feat1), knn2, knn1, ".pos1", ".mlp1"), act=False)
    f3 = fg(pc1, pc2, f1, f2, knn1, knn2, ".pos2", ".mlp2")
    return f1, f2, f3


def _idw(q_t, c_t, feat_t, knn_impl):
    """Inverse-distance 3-NN interpolation shared by PointWarping / UpsampleFlow
    (pointconv_util.py:2131-2139, 2164-2171); all [B,*,C] point-major."""
    B, N, C = q_t.shape
    idx = knn_point(3, c_t, q_t, knn_impl)
    rel = index_points_group(c_t, idx) - q_t.view(B, N, 1, C)
    dist = torch.norm(rel, dim=3).clamp(min=1e-10)
    norm = torch.sum(1.0 / dist, dim=2, keepdim=True)
    weight = (1.0 / dist) / norm
    return torch.sum(weight.view(B, N, 3, 1) * index_points_group(feat_t, idx), dim=2)


def point_warping(xyz1, xyz2, flow1=None, knn_impl="c"):
    """PointWarping.forward, pointconv_util.py:2116-2142."""
    if flow1 is None:
        return xyz2
    x12 = (xyz1 + flow1).permute(0, 2, 1)
    x2 = xyz2.permute(0, 2, 1)
    flow2 = _idw(x2, x12, flow1.permute(0, 2, 1), knn_impl)
    return (x2 - flow2).permute(0, 2, 1)


def upsample_flow(xyz, sparse_xyz, sparse_flow, knn_impl="c"):
    """UpsampleFlow.forward, pointconv_util.py:2154-2172."""
    return _idw(xyz.permute(0, 2, 1), sparse_xyz.permute(0, 2, 1), sparse_flow.permute(0, 2, 1), knn_impl).permute(0, 2, 1)


def scene_flow_estimator_residual(sd, prefix, xyz, feats, cost_volume, flow=None, neighbors=9, clamp=(-200, 200),
                                  training=False, knn_impl="c"):
    """SceneFlowEstimatorResidual.forward, pointconv_util.py:2236-2256."""
    x = torch.cat([feats, cost_volume], dim=1)
    i = 0
    while f"{prefix}.pointconv_list.{i}.linear.weight" in sd:
        x = pointconv(sd, f"{prefix}.pointconv_list.{i}", neighbors, xyz, x, bn=True, training=training, knn_impl=knn_impl)
        i += 1
    i = 0
    while f"{prefix}.mlp_convs.{i}.composed_module.0.weight" in sd:
        x = conv1d(sd, f"{prefix}.mlp_convs.{i}", x)
        i += 1
    local = conv1d(sd, prefix + ".fc", x, act=False).clamp(clamp[0], clamp[1])
    return x, (local if flow is None else local + flow)


# ------------------------------------------------------------------ whole model (a20)
def bid_pointconv_forward(sd, xyz1, xyz2, color1, color2, knn_impl="c"):
    """PointConvBidirection.forward, models_bid_pointconv.py:74-207 (eval mode)."""
    flow_nei, feat_nei = 32, 16
    c1 = lambda name, x: conv1d(sd, name, x)
    up = lambda a, b, c: upsample_flow(a, b, c, knn_impl)
    pc = {1: {}, 2: {}}
    f = {1: {}, 2: {}}
    fps = {1: [], 2: []}
    xyz = {1: xyz1, 2: xyz2}
    col = {1: color1, 2: color2}
    for s in (1, 2):
        pc[s][0] = xyz[s].permute(0, 2, 1)
        f[s]["l0"] = c1("level0_1", c1("level0", col[s].permute(0, 2, 1)))
        f[s]["l0_1"] = c1("level0_2", f[s]["l0"])
        npts = {1: 2048, 2: 512, 3: 256}
        prev = f[s]["l0_1"]
        for lv in (1, 2, 3):
            pc[s][lv], x, idx = pointconvd(sd, f"level{lv}", npts[lv], feat_nei, pc[s][lv - 1], prev, knn_impl)
            fps[s].append(idx)
            f[s][f"l{lv}"] = c1(f"level{lv}_0", x)
            f[s][f"l{lv}_n"] = c1(f"level{lv}_1", f[s][f"l{lv}"])
            prev = f[s][f"l{lv}_n"]
        pc[s][4], x4, _ = pointconvd(sd, "level4", 64, feat_nei, pc[s][3], prev, knn_impl)
        f[s]["l4_3"] = c1("deconv4_3", up(pc[s][3], pc[s][4], x4))

    cf = {s: torch.cat([f[s]["l3"], f[s]["l4_3"]], dim=1) for s in (1, 2)}
    n1, n2, cross3 = cross_layer_light(sd, "cross3", flow_nei, pc[1][3], pc[2][3], cf[1], cf[2], knn_impl)
    feat3, flow3 = scene_flow_estimator_residual(sd, "flow3", pc[1][3], f[1]["l3"], cross3, knn_impl=knn_impl)
    flows, crosses, feat_prev = [flow3], [cross3], feat3
    extra = {1: [], 2: []}
    deconv = {2: "deconv3_2", 1: "deconv2_1", 0: "deconv1_0"}
    for lv in (2, 1, 0):
        new = {1: n1, 2: n2}
        fl = {}
        for s in (1, 2):
            fl[s] = c1(deconv[lv], up(pc[s][lv], pc[s][lv + 1], new[s]))
            extra[s].append(fl[s])
        base = {s: (f[s][f"l{lv}"] if lv > 0 else f[s]["l0"]) for s in (1, 2)}
        cf = {s: torch.cat([base[s], fl[s]], dim=1) for s in (1, 2)}
        up_flow = up(pc[1][lv], pc[1][lv + 1], 1.0 * flows[-1])
        pc2_warp = point_warping(pc[1][lv], pc[2][lv], up_flow, knn_impl)
        n1, n2, cr = cross_layer_light(sd, f"cross{lv}", flow_nei, pc[1][lv], pc2_warp, cf[1], cf[2], knn_impl)
        feat_up = up(pc[1][lv], pc[1][lv + 1], feat_prev)
        feat_prev, fl_new = scene_flow_estimator_residual(sd, f"flow{lv}", pc[1][lv], torch.cat([base[1], feat_up], dim=1),
                                                          cr, up_flow, knn_impl=knn_impl)
        flows.append(fl_new)
        crosses.append(cr)
    flows, crosses = flows[::-1], crosses[::-1]
    feats = {s: [f[s]["l0_1"], f[s]["l1_n"], f[s]["l2_n"], f[s]["l3_n"]] + extra[s] for s in (1, 2)}
    pcs = {s: [pc[s][0], pc[s][1], pc[s][2], pc[s][3]] for s in (1, 2)}
    return flows, fps[1], fps[2], pcs[1], pcs[2], feats[1], feats[2], crosses


def multi_scale_loss(pred_flows: Sequence[torch.Tensor], gt_flow, fps_idxs, alpha=(0.02, 0.04, 0.08, 0.16)):
    """multiScaleLoss, loss_functions.py:6-25."""
    num_scale = len(pred_flows)
    offset = len(fps_idxs) - num_scale + 1
    gts = [gt_flow]
    for idx in fps_idxs:
        gts.append(index_points_gather(gts[-1], idx))
    total = torch.zeros(1)
    for i in range(num_scale):
        diff = pred_flows[i].permute(0, 2, 1) - gts[i + offset]
        total += alpha[i] * torch.norm(diff, dim=2).sum(dim=1).mean()
    return total


def loss_fn_kd_2(outputs, fps_idxs, gt_flow, teacher_flow0, gamma):
    """loss_functions.py:27-36: gamma * MS(student, teacher flow0) + (1-gamma) * MS(student, gt).
    ``teacher_flow0`` is the teacher's finest flow [B,3,N] (the reference indexes teacher_outputs[0])."""
    t0 = teacher_flow0.permute(0, 2, 1)
    return gamma * multi_scale_loss(outputs, t0, fps_idxs) + (1 - gamma) * multi_scale_loss(outputs, gt_flow, fps_idxs)


def bidirection_loss_ht(outputs, feat1s, feat2s, fps_idxs1, gt_flow, teacher_flow0, t_feat1s, t_feat2s, gamma, beta, layer=0):
    """biDirection_loss_ht, loss_functions.py:83-96."""
    t0 = teacher_flow0.permute(0, 2, 1)
    loss1 = multi_scale_loss(outputs, t0, fps_idxs1)
    loss2 = multi_scale_loss(outputs, gt_flow, fps_idxs1)
    src = ((feat1s[layer] - t_feat1s[layer]) ** 2) / 2
    tgt = ((feat2s[layer] - t_feat2s[layer]) ** 2) / 2
    return beta * (gamma * loss1 + (1 - gamma) * loss2) + (1 - beta) * (0.5 * src.sum() + 0.5 * tgt.sum())


def cross_bidirection_loss_ht(outputs, feat1s, fps_idxs1, gt_flow, teacher_flow0, t_feat1s, t_feat2s, gamma, beta, layer=(2, 3)):
    """cross_biDirection_loss_ht, loss_functions.py:201-219 (student feature against cat(teacher feat1, feat2):
    raises unless the student's hint features have twice the teacher's channels, like the reference)."""
    t0 = teacher_flow0.permute(0, 2, 1)
    loss1 = multi_scale_loss(outputs, t0, fps_idxs1)
    loss2 = multi_scale_loss(outputs, gt_flow, fps_idxs1)
    hint = torch.zeros(1)
    for each in layer:
        t_feats = torch.cat([t_feat1s[each], t_feat2s[each]], dim=1)
        hint = hint + ((feat1s[each] - t_feats) ** 2).sum() / 2
    return beta * (gamma * loss1 + (1 - gamma) * loss2) + (1 - beta) * hint
