"""Mirror of the reference's evaluation_utils.py (evaluate_3d, evaluate_2d; :17-50) and of the 2-D flow construction
of utils/geometry.py:6-65, computed by ONE CUDA kernel (csrc/metrics.cu) where the tensors already are.

The reference copies pc1, pc2, the ground truth and the prediction to the host and evaluates with numpy after every
batch (evaluate_bid_pointconv.py:128-145; its ``np.float`` also raises on numpy >= 1.24).  Here:

  * ``scene_flow_metrics(pc1, pred_flow, gt_flow, calib=None)`` -> float32[6] DEVICE tensor
    (EPE3D, Acc3DS, Acc3DR, Outliers3D, EPE2D, Acc2D), no synchronisation;
  * ``MetricMeter`` accumulates batches on the device (AverageMeter semantics: plain mean over batches, as
    evaluate_bid_pointconv.py:130-146) and synchronises once, in ``result()``;
  * ``evaluate_3d`` / ``evaluate_2d`` keep the reference's names, argument order and return tuples for callers that
    still hand in numpy arrays (they are uploaded, evaluated on the GPU and returned as Python floats).
There is no CPU implementation: a CPU tensor without a CUDA device raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops  # noqa: F401  (registers torch.ops.kdpc)

K = torch.ops.kdpc
NAMES = ("EPE3D", "ACC3DS", "ACC3DR", "Outliers3D", "EPE2D", "ACC2D")


def scene_flow_metrics(pc1: Optional[torch.Tensor], pred_flow: torch.Tensor, gt_flow: torch.Tensor,
                       calib: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pc1, gt_flow: [B,N,3]; pred_flow: the model's [B,3,N] output or a point-major [B,N,3] tensor
    (when N == 3 pass [B,3,3] channel-major, as the model produces it).  calib: None or [B,6]."""
    gt = gt_flow.contiguous()
    point_major = pred_flow.dim() == 3 and tuple(pred_flow.shape) == tuple(gt.shape) and gt.shape[1] != 3
    if not point_major and pred_flow.dim() == 3 and pred_flow.shape[1] != 3:
        raise ValueError("kdpc: pred_flow must be [B,3,N] or [B,N,3]")
    return K.flow_metrics(pred_flow.contiguous(), gt, None if pc1 is None else pc1.contiguous(),
                          None if calib is None else calib.contiguous().float(), point_major)


class MetricMeter:
    """Device-side running mean of the six metrics over batches (the reference's six AverageMeters)."""

    def __init__(self):
        self.sum = None
        self.count = 0

    def update(self, pc1, pred_flow, gt_flow, calib=None) -> torch.Tensor:
        m = scene_flow_metrics(pc1, pred_flow, gt_flow, calib)
        self.sum = m.clone() if self.sum is None else self.sum + m
        self.count += 1
        return m

    def result(self) -> dict:
        if self.sum is None:
            return {k: float("nan") for k in NAMES}
        vals = (self.sum / self.count).tolist()              # the only device->host synchronisation
        return dict(zip(NAMES, vals))


def _to_device(a) -> torch.Tensor:
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32) if not torch.is_tensor(a) else a.float()
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("kdpc: evaluation metrics run on the GPU only (no CPU implementation)")
        t = t.cuda()
    return t.reshape(1, -1, t.shape[-1]) if t.dim() == 2 else t.reshape(t.shape[0], -1, t.shape[-1])


def evaluate_3d(sf_pred, sf_gt) -> Tuple[float, float, float, float]:
    """evaluation_utils.py:17-33: (N,3) or (B,N,3) arrays -> EPE3D, acc3d_strict, acc3d_relax, outlier."""
    p, g = _to_device(sf_pred), _to_device(sf_gt)
    m = K.flow_metrics(p.contiguous(), g.contiguous(), None, None, True).tolist()
    return m[0], m[1], m[2], m[3]


def evaluate_2d(flow_pred, flow_gt) -> Tuple[float, float]:
    """evaluation_utils.py:36-50 on already projected 2-D flows ((N,2) or (B,N,2) arrays): the few elementwise torch
    ops of the reference formula on the device (the fused kernel projects by itself - use scene_flow_metrics)."""
    p, g = _to_device(flow_pred), _to_device(flow_gt)
    epe = torch.linalg.vector_norm(g - p, dim=-1)
    rel = epe / (torch.linalg.vector_norm(g, dim=-1) + 1e-5)
    acc = ((epe < 3.0) | (rel < 0.05)).double().mean()
    return float(epe.double().mean()), float(acc)
