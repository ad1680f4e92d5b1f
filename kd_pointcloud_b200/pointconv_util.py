"""Drop-in for the hot subset of the reference's ``pointconv_util.py`` / ``pointconv_util2.py``.

Same class / function names, constructor signatures, forward signatures, return layouts and
``state_dict`` keys, so ``models_bid_pointconv.py``, ``models_bid_lighttoken_res.py``
and ``distilTrain.py`` import it unchanged.  The forwards do NOT follow the reference's op
chains: they keep activations point-major ([B,N,C], returned to callers as permuted [B,C,N]
views — which is what the reference's own layers return, pointconv_util.py:255,444) and call
the fused sm_100a kernels in libkdpc.  There is no CPU path.

Reference anchors per class are given in the docstrings.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as KF
from . import pointnet2_utils  # noqa: F401  (re-exported like the reference does)
from .functional import cm, pm, knn_idx, knn_point, square_distance  # noqa: F401

LEAKY_RATE = 0.1
use_bn = False

K = torch.ops.kdpc
FUSED_COSTVOL = True      # tests flip this to compare the fused kernel with the op chain


def _act(use_leaky: bool) -> nn.Module:
    return nn.LeakyReLU(LEAKY_RATE, inplace=True) if use_leaky else nn.ReLU(inplace=True)


def _slope(m: nn.Module) -> float:
    return float(m.negative_slope) if isinstance(m, nn.LeakyReLU) else 0.0


def _is_pointwise(conv) -> bool:
    one = lambda v: all(int(x) == 1 for x in (v if isinstance(v, (tuple, list)) else (v,)))
    zero = lambda v: all(int(x) == 0 for x in (v if isinstance(v, (tuple, list)) else (v,)))
    return one(conv.kernel_size) and one(conv.stride) and zero(conv.padding) and conv.groups == 1


def _linear_pm(conv: nn.Module, x: torch.Tensor, norm: Optional[nn.Module] = None, act: Optional[nn.Module] = None,
               clamp=None, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A 1x1 nn.Conv1d / nn.Conv2d / nn.Linear applied along the LAST axis of a point-major tensor,
    followed by an optional BatchNorm (over all leading axes), activation, clamp and residual add.
    Inference: ONE fused tcgen05 kernel (functional.fused_linear).  When a gradient is needed or
    the BatchNorm is in training mode: the same math from torch ops (autograd).
    ``out`` (inference only, see ``_concat_free``): a row-strided view the fused kernel writes into."""
    w = conv.weight
    bn = None if (norm is None or isinstance(norm, nn.Identity)) else norm
    slope = 1.0 if act is None else _slope(act)
    if KF.fused_linear_available(x, w, conv.bias, bn):
        return KF.fused_linear(x, w, conv.bias, bn, slope, clamp, residual, out=out)
    if out is not None:
        raise RuntimeError("kdpc: an output view is only supported on the fused inference path")
    w2d = w.reshape(w.shape[0], -1)
    if KF.linear_tc_autograd_available(x, w2d):
        if bn is None and act is not None and 0.0 <= slope < 1.0:
            # tcgen05 forward with the activation in its epilogue + tcgen05 dX / dW (training path)
            y = KF.linear_tc_autograd(x, w2d, conv.bias, slope)
            act = None
        else:
            y = KF.linear_tc_autograd(x, w2d, conv.bias)
    elif KF.linear_small_autograd_available(x, w2d):
        y = KF.linear_small_autograd(x, w2d, conv.bias)     # 3 -> D / D -> 3 layers: SIMT forward + dX, tcgen05 dW
    else:
        y = F.linear(x, w2d, conv.bias)
    if bn is not None:
        y = bn(y.reshape(-1, y.shape[-1])).view(y.shape)
    if act is not None:
        y = F.leaky_relu(y, slope) if slope != 0.0 else F.relu(y)
    if clamp is not None:
        y = y.clamp(clamp[0], clamp[1])
    if residual is not None:
        y = y + residual
    return y


class _ComposedConv(nn.Module):
    """Conv + (BatchNorm | Identity) + activation, registered as ``composed_module`` exactly like
    the reference's Conv1d / Conv2d (pointconv_util.py:20-54)."""

    def forward(self, x):
        conv, norm, act = self.composed_module[0], self.composed_module[1], self.composed_module[2]
        if not _is_pointwise(conv) or not isinstance(norm, nn.Identity):
            return self.composed_module(x)
        # 1x1 conv == GEMM over channels; keep the result point-major and hand back a view
        perm_in = (0, 2, 1) if x.dim() == 3 else (0, 3, 2, 1)
        return _linear_pm(conv, x.permute(*perm_in), None, act).permute(*perm_in)

    def forward_pm(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x [..., Cin] -> [..., Cout] (channels last)."""
        conv, norm, act = self.composed_module[0], self.composed_module[1], self.composed_module[2]
        if not _is_pointwise(conv):
            raise NotImplementedError("forward_pm needs a 1x1 convolution")
        return _linear_pm(conv, x, norm, act, out=out)


class Conv1d(_ComposedConv):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, use_leaky=True, bn=use_bn):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.composed_module = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=True),
            nn.BatchNorm1d(out_channels) if bn else nn.Identity(),
            _act(use_leaky))


class Conv2d(_ComposedConv):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, use_leaky=True, bn=use_bn,
                 bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.composed_module = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias),
            nn.BatchNorm2d(out_channels) if bn else nn.Identity(),
            _act(use_leaky))


class ConvBNReLU(nn.Module):
    """pointconv_util3.py:69-79 (depthwise 1x1 conv + ReLU; despite the name there is no BN)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, affine=True):
        super().__init__()
        self.op = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                      groups=in_channels, bias=False),
            nn.ReLU(inplace=False))

    def forward(self, x):
        return self.op(x)


class BottleNeck(nn.Module):
    """pointconv_util3.py:51-67.  Imported (not used on the hot path) by models_bid_pointconv.py:7;
    the reference's own pointconv_util.py lacks it."""

    def __init__(self, in_channels, mid_channel, out_channel, kernel_size=1, stride=1, padding=0, use_leaky=True,
                 bn=use_bn):
        super().__init__()
        self.bottleneck = nn.Sequential(
            nn.Conv1d(in_channels, mid_channel, kernel_size=1),
            nn.Conv1d(mid_channel, mid_channel, kernel_size=3, padding=1, bias=False),
            nn.Conv1d(mid_channel, out_channel, kernel_size=1))
        self.depthwiseConv = ConvBNReLU(in_channels, out_channel)
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(self.bottleneck(x) + x + self.depthwiseConv(x))


# ------------------------------------------------------------------------------- a8, a9
def index_points_gather(points: torch.Tensor, fps_idx: torch.Tensor) -> torch.Tensor:
    """points [B,N,C], fps_idx [B,S] -> [B,S,C]   (pointconv_util.py:109-120)."""
    return KF.gather_rows(points, fps_idx)


def index_points_group(points: torch.Tensor, knn_idx_: torch.Tensor) -> torch.Tensor:
    """points [B,N,C], knn_idx [B,S,K] -> [B,S,K,C]   (pointconv_util.py:122-133)."""
    return KF.gather_rows(points, knn_idx_)


def group(nsample, xyz, points):
    """pointconv_util.py:135-157: returns (new_points [B,N,K,3+D], grouped_xyz_norm [B,N,K,3])."""
    return group_query(nsample, xyz, xyz, points)


def group_query(nsample, s_xyz, xyz, s_points):
    """pointconv_util.py:159-182: neighbours of ``xyz`` (queries) among ``s_xyz`` (support)."""
    idx = knn_idx(nsample, s_xyz, xyz)
    new_points = KF.group_concat(s_xyz, xyz, s_points, idx)
    return new_points, new_points[..., :3]


# ----------------------------------------------------------------------------------- a10
class WeightNet(nn.Module):
    """pointconv_util.py:184-215.  ReLU after every layer including the last; ``mlp_bns`` are
    built (and kept in the state_dict) even though ``bn=False`` never uses them."""

    def __init__(self, in_channel, out_channel, hidden_unit=[8, 8], bn=use_bn):
        super().__init__()
        self.bn = bn
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        widths = [in_channel] + list(hidden_unit or []) + [out_channel]
        for cin, cout in zip(widths[:-1], widths[1:]):
            self.mlp_convs.append(nn.Conv2d(cin, cout, 1))
            self.mlp_bns.append(nn.BatchNorm2d(cout))

    def _fusable(self, x: torch.Tensor) -> bool:
        c = self.mlp_convs
        return (not self.bn and len(c) == 3 and c[0].in_channels == 3 and c[0].out_channels == 8
                and c[1].out_channels == 8 and c[2].out_channels in (4, 8, 16, 32, 48)
                and not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))))

    def forward_pm(self, rel: torch.Tensor, coords_need_grad: bool = True) -> torch.Tensor:
        """rel [..., W>=3] whose first 3 channels are the localized xyz -> [..., out_channel].
        ``coords_need_grad=False`` (the coordinates are constants, as in every PointConv of the models): the training
        path does not route a gradient back into ``rel`` through the WeightNet."""
        if self._fusable(rel):
            c = self.mlp_convs
            return K.weightnet(rel.contiguous(), c[0].weight.reshape(8, 3), c[0].bias, c[1].weight.reshape(8, 8),
                               c[1].bias, c[2].weight.reshape(-1, 8), c[2].bias)
        c = self.mlp_convs
        if (KF.USE_FUSED_WEIGHTNET_GRAD and not self.bn and rel.is_cuda and len(c) == 3 and c[0].in_channels == 3 and c[0].out_channels == 8
                and c[1].out_channels == 8 and c[2].out_channels in (8, 16) and all(_is_pointwise(m) and m.bias is not None for m in c)):
            return KF.weightnet_autograd(rel if coords_need_grad else rel.detach(), c)   # training: fused forward + one-kernel backward
        w = rel[..., :self.mlp_convs[0].in_channels]
        for i, conv in enumerate(self.mlp_convs):
            w = _linear_pm(conv, w)
            if self.bn:
                w = self.mlp_bns[i](w.permute(0, 3, 2, 1)).permute(0, 3, 2, 1)      # rel is [B,S,K,C]
            w = F.relu(w)
        return w

    def forward(self, localized_xyz: torch.Tensor) -> torch.Tensor:
        # reference layout: [B,C,K,N] -> [B,W,K,N]
        return self.forward_pm(localized_xyz.permute(0, 3, 2, 1)).permute(0, 3, 2, 1)


# ------------------------------------------------------------------------------ a11, a12
class _PointConvBase(nn.Module):
    def _init_common(self, nsample, in_channel, out_channel, weightnet, bn, use_leaky):
        self.bn = bn
        self.nsample = nsample
        self.weightnet = WeightNet(3, weightnet)
        self.linear = nn.Linear(weightnet * in_channel, out_channel)
        if bn:
            self.bn_linear = nn.BatchNorm1d(out_channel)
        self.relu = _act(use_leaky)

    def _contract(self, s_xyz, q_xyz, s_points, idx) -> torch.Tensor:
        """grouping + WeightNet + sum over K + Linear (+BN) + activation; everything point-major.
        s_xyz [B,N,3] support, q_xyz [B,S,3] queries, s_points [B,N,D] -> [B,S,Cout]."""
        bn = self.bn_linear if self.bn else None
        if KF.fused_pointconv_available(self.weightnet, self.linear, bn, idx.shape[2], s_points):
            return KF.fused_pointconv(s_xyz, q_xyz, s_points, idx, self.weightnet, self.linear, bn, _slope(self.relu))
        grouped = KF.group_concat(s_xyz, q_xyz, s_points, idx)            # [B,S,K,3+D]
        wn = self.weightnet.forward_pm(grouped, s_xyz.requires_grad or q_xyz.requires_grad)   # [B,S,K,W]
        agg = KF.pointconv_agg(grouped, wn)                               # [B,S,(3+D)*W]  (c-major)
        # Linear (+ BatchNorm1d over B and S) + activation
        return _linear_pm(self.linear, agg, self.bn_linear if self.bn else None, self.relu)


class PointConv(_PointConvBase):
    """pointconv_util.py:217-258.  xyz [B,3,N], points [B,D,N] -> [B,Cout,N]."""

    def __init__(self, nsample, in_channel, out_channel, weightnet=16, bn=use_bn, use_leaky=True):
        super().__init__()
        self._init_common(nsample, in_channel, out_channel, weightnet, bn, use_leaky)

    def forward_pm(self, xyz_pm: torch.Tensor, points_pm: torch.Tensor) -> torch.Tensor:
        idx = knn_idx(self.nsample, xyz_pm, xyz_pm)
        return self._contract(xyz_pm, xyz_pm, points_pm, idx)

    def forward(self, xyz, points):
        return cm(self.forward_pm(pm(xyz), pm(points)))


class PointConvD(_PointConvBase):
    """pointconv_util.py:401-446.  Returns (new_xyz [B,3,S], feats [B,Cout,S], fps_idx int32 [B,S])."""

    def __init__(self, npoint, nsample, in_channel, out_channel, weightnet=16, bn=use_bn, use_leaky=True):
        super().__init__()
        self.npoint = npoint
        self._init_common(nsample, in_channel, out_channel, weightnet, bn, use_leaky)

    def forward_pm(self, xyz_pm, points_pm, sampled=None):
        """``sampled`` = (fps_idx, new_xyz) when the caller already ran the sampling (flownet.py computes the whole
        FPS pyramid up front on a side stream: it depends on coordinates only)."""
        if sampled is None:
            fps_idx = KF.furthest_point_sample(xyz_pm, self.npoint)
            new_xyz = KF.gather_rows(xyz_pm, fps_idx)                     # [B,S,3]
        else:
            fps_idx, new_xyz = sampled
        idx = knn_idx(self.nsample, xyz_pm, new_xyz)
        return new_xyz, self._contract(xyz_pm, new_xyz, points_pm, idx), fps_idx

    def forward(self, xyz, points):
        new_xyz, feats, fps_idx = self.forward_pm(pm(xyz), pm(points))
        return cm(new_xyz), cm(feats), fps_idx


class PointConvWeight(PointConvD):
    """pointconv_util2.py:434-481: textually PointConvD under another name (models_bid_lighttoken_weight48.py)."""


# ----------------------------------------------------------------------------------- a13
class CrossLayerLight(nn.Module):
    """Bidirectional cost volume, pointconv_util.py:1791-1868."""

    def __init__(self, nsample, in_channel, mlp1, mlp2, bn=use_bn, use_leaky=True):
        super().__init__()
        self.nsample = nsample
        self.bn = bn
        self.pos1 = nn.Conv2d(3, mlp1[0], 1)
        self.mlp1 = nn.ModuleList()
        self.cross_t11 = nn.Conv1d(in_channel, mlp1[0], 1)
        self.cross_t22 = nn.Conv1d(in_channel, mlp1[0], 1)
        self.bias1 = nn.Parameter(torch.randn((1, mlp1[0], 1, 1)), requires_grad=True)   # unused in cross()
        self.bn1 = nn.BatchNorm2d(mlp1[0]) if bn else nn.Identity()
        for cin, cout in zip(mlp1[:-1], mlp1[1:]):
            self.mlp1.append(Conv2d(cin, cout, bn=bn, use_leaky=use_leaky))

        self.mlp2 = True if mlp2 is not None else False
        if mlp2 is not None:
            self.cross_t1 = nn.Conv1d(mlp1[-1], mlp2[0], 1)
            self.cross_t2 = nn.Conv1d(mlp1[-1], mlp2[0], 1)
            self.pos2 = nn.Conv2d(3, mlp2[0], 1)
            self.bias2 = nn.Parameter(torch.randn((1, mlp2[0], 1, 1)), requires_grad=True)  # unused
            self.bn2 = nn.BatchNorm2d(mlp2[0]) if bn else nn.Identity()
            self.mlp2 = nn.ModuleList()
            for cin, cout in zip(mlp2[:-1], mlp2[1:]):
                self.mlp2.append(Conv2d(cin, cout, bn=bn, use_leaky=use_leaky))
        self.relu = _act(use_leaky)

    def cross_pm(self, xyz1, xyz2, points1, points2, pos, mlp, bn, idx=None) -> torch.Tensor:
        """All point-major: xyz1 [B,N1,3] queries, xyz2 [B,N2,3], points1 [B,N1,D], points2 [B,N2,D]
        -> [B,N1,D'].   relu(bn(p2[idx] + p1 + pos(xyz2[idx]-xyz1))) -> mlp -> max over K.
        ``idx`` int32 [B,N1,K]: the neighbourhood when the caller builds it itself (CrossLayerLightFG)."""
        if idx is None:
            idx = knn_idx(self.nsample, xyz2, xyz1)
        D = points1.shape[2]
        needs_grad = torch.is_grad_enabled() and any(
            t.requires_grad for t in (xyz1, xyz2, points1, points2, pos.weight, pos.bias))
        if (FUSED_COSTVOL and isinstance(bn, nn.Identity) and not needs_grad and idx.shape[2] == 32 and D % 8 == 0 and D <= 256
                and points2.shape[2] == D and len(mlp) == 1 and _is_pointwise(mlp[0].composed_module[0])
                and isinstance(mlp[0].composed_module[1], nn.Identity) and mlp[0].out_channels <= 256
                and KF.fused_linear_available(points1, mlp[0].composed_module[0].weight, mlp[0].composed_module[0].bias, None)):
            return KF.fused_costvol(xyz1, xyz2, points1, points2, idx, pos, _slope(self.relu), mlp[0].composed_module[0],
                                    _slope(mlp[0].composed_module[2]))
        if (needs_grad and FUSED_COSTVOL and isinstance(bn, nn.Identity) and points2.shape[2] == D and len(mlp) == 1
                and _is_pointwise(mlp[0].composed_module[0]) and isinstance(mlp[0].composed_module[1], nn.Identity)
                and KF.costvol_autograd_available(points1, idx, mlp[0].composed_module[0], _slope(self.relu))):
            # training, 8192-point level: fused forward + recomputing arg-max backward (csrc/costvol_grad.cu)
            return KF.costvol_autograd(xyz1, xyz2, points1, points2, idx, pos, _slope(self.relu), mlp[0].composed_module[0],
                                       _slope(mlp[0].composed_module[2]))
        if isinstance(bn, nn.Identity) and D % 4 == 0 and points2.shape[2] == D and not needs_grad:
            x = K.costvol_pre(xyz1.contiguous(), xyz2.contiguous(), points1.contiguous(), points2.contiguous(), idx,
                              pos.weight.reshape(D, 3), pos.bias, _slope(self.relu))
        else:
            B, N1, _ = xyz1.shape
            rel = KF.group_concat(xyz2, xyz1, None, idx)                              # [B,N1,K,3]
            x = KF.gather_rows(points2, idx) + points1.view(B, N1, 1, D) + _linear_pm(pos, rel)
            if not isinstance(bn, nn.Identity):
                x = bn(x.permute(0, 3, 2, 1)).permute(0, 3, 2, 1)
            x = self.relu(x)
        for conv in mlp:
            x = conv.forward_pm(x)
        if torch.is_grad_enabled() and x.requires_grad:
            return torch.max(x, dim=2)[0]
        return K.max_over_k(x.contiguous())[0]

    def cross(self, xyz1, xyz2, points1, points2, pos, mlp, bn):
        """Reference signature (pointconv_util.py:1826): channel-major in, [B,D',N1] out."""
        return cm(self.cross_pm(pm(xyz1), pm(xyz2), pm(points1), pm(points2), pos, mlp, bn))

    def forward_pm(self, pc1, pc2, feat1, feat2, feat_both=None):
        """``feat_both``: optionally the [2B,N,C] tensor whose halves are feat1 and feat2 (the encoder keeps both clouds
        in one batch): cross_t11 / cross_t22 then run once over 2B clouds instead of once per cloud (same values)."""
        if feat_both is not None and feat_both.shape[0] == 2 * feat1.shape[0] and feat1.shape == feat2.shape:
            B = feat1.shape[0]
            t11, t22 = _linear_pm(self.cross_t11, feat_both), _linear_pm(self.cross_t22, feat_both)
            t11_1, t11_2, t22_1, t22_2 = t11[:B], t11[B:], t22[:B], t22[B:]
        else:
            t11_1, t22_2 = _linear_pm(self.cross_t11, feat1), _linear_pm(self.cross_t22, feat2)
            t11_2, t22_1 = _linear_pm(self.cross_t11, feat2), _linear_pm(self.cross_t22, feat1)
        a = self.cross_pm(pc1, pc2, t11_1, t22_2, self.pos1, self.mlp1, self.bn1)
        b = self.cross_pm(pc2, pc1, t11_2, t22_1, self.pos1, self.mlp1, self.bn1)
        if self.mlp2 is False:
            return a, b
        if a.shape == b.shape and KF.concat_free(a, b):
            # inference: both directions into the halves of ONE [2B,N,C] tensor - the caller's cat([a, b], dim=0)
            # (the next level's upsampling runs over both clouds at once) is then functional.joined(a, b), no copy
            ab = a.new_empty((2 * a.shape[0],) + tuple(a.shape[1:-1]) + (self.cross_t1.out_channels,))
            a = _linear_pm(self.cross_t1, a, out=ab[:a.shape[0]])
            b = _linear_pm(self.cross_t2, b, out=ab[a.shape[0]:])
        else:
            a = _linear_pm(self.cross_t1, a)
            b = _linear_pm(self.cross_t2, b)
        c = self.cross_pm(pc1, pc2, a, b, self.pos2, self.mlp2, self.bn2)
        return a, b, c

    def forward(self, pc1, pc2, feat1, feat2):
        return tuple(cm(t) for t in self.forward_pm(pm(pc1), pm(pc2), pm(feat1), pm(feat2)))


class NoCrossLayerLight(CrossLayerLight):
    """pointconv_util.py:1276-1331 (pointconv_util2.py:1197): ONE cost-volume direction, cross_t1/cross_t2 + pos + mlp.
    ``bn`` is used by truthiness exactly like the reference (models_bid_no_cross.py:26 passes a list in its place)."""

    def __init__(self, nsample, in_channel, mlp1, bn=use_bn, use_leaky=True, output_clue=False):
        nn.Module.__init__(self)
        self.nsample = nsample
        self.output_clue = output_clue
        self.mlp1_convs = nn.ModuleList()
        if bn:
            self.mlp1_bns = nn.ModuleList()
        self.cross_t1 = nn.Conv1d(in_channel, mlp1[0], 1)
        self.cross_t2 = nn.Conv1d(in_channel, mlp1[0], 1)
        self.pos = nn.Conv2d(3, mlp1[0], 1)
        self.bias = nn.Parameter(torch.randn((1, mlp1[0], 1, 1)), requires_grad=True)       # unused in cross()
        self.bn = nn.BatchNorm2d(mlp1[0]) if bn else nn.Identity()
        self.mlp = nn.ModuleList()
        for cin, cout in zip(mlp1[:-1], mlp1[1:]):
            self.mlp.append(Conv2d(cin, cout, bn=bn, use_leaky=use_leaky))
        self.relu = _act(use_leaky)

    def forward_pm(self, pc1, pc2, feat1, feat2):
        return self.cross_pm(pc1, pc2, _linear_pm(self.cross_t1, feat1), _linear_pm(self.cross_t2, feat2), self.pos, self.mlp, self.bn)

    def forward(self, pc1, pc2, feat1, feat2):
        return cm(self.forward_pm(pm(pc1), pm(pc2), pm(feat1), pm(feat2)))


class CrossLayerLightFG(CrossLayerLight):
    """pointconv_util.py:1871-1957: the cost volume over a neighbourhood made of nsample/2 nearest neighbours in a
    FEATURE space (knn1 / knn2, any channel count: csrc/knn_feat.cu) followed by nsample/2 nearest in xyz; all three
    cross() calls are followed by cross_t1 / cross_t2 as in the reference's forward (:1944-1957)."""

    def cross_fg_pm(self, xyz1, xyz2, points1, points2, knn1, knn2, pos, mlp, bn, nsample=None):
        half = (self.nsample if nsample is None else nsample) // 2
        idx = torch.cat([knn_idx(half, knn2, knn1), knn_idx(half, xyz2, xyz1)], dim=2).contiguous()
        return self.cross_pm(xyz1, xyz2, points1, points2, pos, mlp, bn, idx=idx)

    def cross(self, xyz1, xyz2, points1, points2, knn1, knn2, pos, mlp, bn, nsample=None):
        return cm(self.cross_fg_pm(pm(xyz1), pm(xyz2), pm(points1), pm(points2), pm(knn1), pm(knn2), pos, mlp, bn, nsample))

    def forward_pm(self, pc1, pc2, feat1, feat2, knn1, knn2):
        a = self.cross_fg_pm(pc1, pc2, _linear_pm(self.cross_t11, feat1), _linear_pm(self.cross_t22, feat2), knn1, knn2,
                             self.pos1, self.mlp1, self.bn1)
        a = _linear_pm(self.cross_t1, a)
        b = self.cross_fg_pm(pc2, pc1, _linear_pm(self.cross_t11, feat2), _linear_pm(self.cross_t22, feat1), knn2, knn1,
                             self.pos1, self.mlp1, self.bn1)
        b = _linear_pm(self.cross_t2, b)
        c = self.cross_fg_pm(pc1, pc2, a, b, knn1, knn2, self.pos2, self.mlp2, self.bn2)
        return a, b, c

    def forward(self, pc1, pc2, feat1, feat2, knn1, knn2):
        return tuple(cm(t) for t in self.forward_pm(pm(pc1), pm(pc2), pm(feat1), pm(feat2), pm(knn1), pm(knn2)))


# ------------------------------------------------------------------------------ a14, a15
def _interp(q_xyz, c_xyz, feat, idx=None):
    """3-NN inverse-distance interpolation of feat (given at c_xyz) onto q_xyz; point-major.
    ``idx`` (int32 [B,N,3]) may be supplied when the caller already holds knn_idx(3, c_xyz, q_xyz)."""
    if idx is None:
        idx = knn_idx(3, c_xyz, q_xyz)
    if torch.is_grad_enabled() and (c_xyz.requires_grad or q_xyz.requires_grad):
        return KF.interp3_composite(q_xyz, c_xyz, idx, feat)
    return KF.interp3(q_xyz, c_xyz, idx, feat)


class PointWarping(nn.Module):
    """pointconv_util.py:2114-2142: xyz2 - interp(flow1 carried to xyz1 + flow1)."""

    def forward_pm(self, xyz1, xyz2, flow1=None):
        if flow1 is None:
            return xyz2
        carried = xyz1 + flow1
        KF.hint_displaced_copy(carried, xyz1)               # both warped clouds are displaced copies of sorted clouds:
        warped = xyz2 - _interp(xyz2, carried, flow1)       # their kNNs reuse the parents' Morton order (same results)
        KF.hint_displaced_copy(warped, xyz2)
        return warped

    def forward(self, xyz1, xyz2, flow1=None):
        if flow1 is None:
            return xyz2
        return cm(self.forward_pm(pm(xyz1), pm(xyz2), pm(flow1)))


class UpsampleFlow(nn.Module):
    """pointconv_util.py:2153-2172: xyz [B,3,N] dense, sparse_xyz [B,3,S], sparse_flow [B,C,S] -> [B,C,N]."""

    def forward_pm(self, xyz, sparse_xyz, sparse_flow, idx=None):
        return _interp(xyz, sparse_xyz, sparse_flow, idx)

    def forward(self, xyz, sparse_xyz, sparse_flow):
        return cm(self.forward_pm(pm(xyz), pm(sparse_xyz), pm(sparse_flow)))


# ----------------------------------------------------------------------------------- a16
class SceneFlowEstimatorResidual(nn.Module):
    """pointconv_util.py:2215-2256."""

    def __init__(self, feat_ch, cost_ch, flow_ch=3, channels=[128, 128], mlp=[128, 64], neighbors=9,
                 clamp=[-200, 200], use_leaky=True, weightnet=16):
        super().__init__()
        self.clamp = clamp
        self.use_leaky = use_leaky
        self.pointconv_list = nn.ModuleList()
        last = feat_ch + cost_ch
        for ch_out in channels:
            self.pointconv_list.append(PointConv(neighbors, last + 3, ch_out, bn=True, use_leaky=True, weightnet=weightnet))
            last = ch_out
        self.mlp_convs = nn.ModuleList()
        for ch_out in mlp:
            self.mlp_convs.append(Conv1d(last, ch_out))
            last = ch_out
        self.fc = nn.Conv1d(last, 3, 1)

    def forward_pm(self, xyz, feats, cost_volume, flow=None):
        """``feats``: one [B,N,C] tensor or a tuple of them (concatenated here together with the cost volume in ONE
        copy instead of the caller's cat followed by this one: same tensor)."""
        parts = [*(feats if isinstance(feats, (tuple, list)) else (feats,)), cost_volume]
        if KF.concat_free(*parts) and len(parts) <= 4 and all(
                t.shape[-1] % 4 == 0 and t.data_ptr() % 16 == 0 and (KF.ops.row_stride(t) or 1) % 4 == 0 for t in parts):
            x = KF.ops.concat_rows(parts)          # one vectorised copy, strided parts welcome (csrc/group.cu)
        else:
            x = torch.cat(parts, dim=2)
        for pointconv in self.pointconv_list:
            x = pointconv.forward_pm(xyz, x)
        for conv in self.mlp_convs:
            x = conv.forward_pm(x)
        # flow = clamp(fc(x)) + up_flow      (pointconv_util.py:2250-2255)
        return x, _linear_pm(self.fc, x, None, None, (self.clamp[0], self.clamp[1]), flow)

    def forward(self, xyz, feats, cost_volume, flow=None):
        x, f = self.forward_pm(pm(xyz), pm(feats), pm(cost_volume), None if flow is None else pm(flow))
        return cm(x), cm(f)


__all__: List[str] = [
    "LEAKY_RATE", "use_bn", "Conv1d", "Conv2d", "ConvBNReLU", "BottleNeck", "square_distance", "knn_point",
    "index_points_gather", "index_points_group", "group", "group_query", "WeightNet", "PointConv", "PointConvD",
    "CrossLayerLight", "CrossLayerLightFG", "NoCrossLayerLight", "PointConvWeight", "PointWarping", "UpsampleFlow", "SceneFlowEstimatorResidual", "pointnet2_utils",
]
