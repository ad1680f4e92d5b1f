"""Drop-in for the reference's ``pointnet2/pointnet2_utils.py`` (same names, argument meaning,
return layouts and ``assert is_contiguous()`` behaviour) on top of libkdpc.

    furthest_point_sample(xyz[B,N,3], npoint)          -> int32 [B,npoint]      (pointnet2_utils.py:10-36)
    gather_operation(features[B,C,N], idx[B,M])        -> [B,C,M]               (:39-73)
    three_nn(unknown[B,n,3], known[B,m,3])             -> (sqrt(d2)[B,n,3], int32 idx)   (:76-105)
    three_interpolate(features[B,C,m], idx, weight)    -> [B,C,n]               (:108-153)
    grouping_operation(features[B,C,N], idx[B,S,K])    -> [B,C,S,K]             (:156-197)
    ball_query(radius, nsample, xyz, new_xyz)          -> int32 [B,npoint,nsample]   (:200-229)
    QueryAndGroup, GroupAll                                                      (:232-290)

Differences, all deliberate: outputs live on the device of the inputs (the reference allocates on
the *current* device); backward passes are deterministic; launch failures raise instead of
``exit(-1)``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as F_

furthest_point_sample = F_.furthest_point_sample
gather_operation = F_.gather_operation
three_nn = F_.three_nn
three_interpolate = F_.three_interpolate
grouping_operation = F_.grouping_operation
ball_query = F_.ball_query


class _ApplyShim:
    """The reference exposes autograd.Function classes whose ``.apply`` is the public op."""

    def __init__(self, fn):
        self.apply = fn


FurthestPointSampling = _ApplyShim(furthest_point_sample)
GatherOperation = _ApplyShim(gather_operation)
ThreeNN = _ApplyShim(three_nn)
ThreeInterpolate = _ApplyShim(three_interpolate)
GroupingOperation = _ApplyShim(grouping_operation)
BallQuery = _ApplyShim(ball_query)


class QueryAndGroup(nn.Module):
    """Ball query + grouping (+ centred xyz), as pointnet2_utils.py:232-268."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        centred = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return centred
        grouped = grouping_operation(features, idx)
        return torch.cat([centred, grouped], dim=1) if self.use_xyz else grouped


class GroupAll(nn.Module):
    """Single group holding every point, as pointnet2_utils.py:271-290."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: Optional[torch.Tensor] = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
