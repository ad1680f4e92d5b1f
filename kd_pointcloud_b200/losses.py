"""Multi-scale flow loss and distillation terms (reference loss_functions.py).

    multiScaleLoss              loss_functions.py:6-25  (copies: models_bid_pointconv.py:545-563)
    loss_fn_kd_2                loss_functions.py:27-36
    biDirection_loss_ht         loss_functions.py:83-96
    cross_biDirection_loss_ht   loss_functions.py:201-219  (the one distilTrain.py:174 calls; it raises
                                for the shipped student because cat(t_feat1,t_feat2) has twice the
                                student's channels — SURVEY 9.  ``hint_mode='first'`` gives the
                                shape-valid sibling used by BASELINE config 4.)
    epe3d                       distilTrain.py:229

The GT pyramid is the chained FPS gather of the ground-truth flow; the per-scale term is
``alpha_i * mean_b sum_n ||pred_i - gt_i||_2``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import functional as KF

scale = 1.0
ALPHA = (0.02, 0.04, 0.08, 0.16)


FUSED = True      # tests flip this to compare the fused kernel with the reference op chain


def _fusable(pred_flows, gt_flow, fps_idxs, hints=()) -> bool:
    ts = list(pred_flows) + [gt_flow] + [h for pair in hints for h in pair]
    return (FUSED and scale == 1.0 and all(t.is_cuda and t.dtype == torch.float32 for t in ts)
            and 1 <= len(pred_flows) <= 4 and len(fps_idxs) == len(pred_flows) - 1 and not gt_flow.requires_grad)


def multiScaleLoss(pred_flows: Sequence[torch.Tensor], gt_flow: torch.Tensor, fps_idxs: Sequence[torch.Tensor],
                   alpha: Sequence[float] = ALPHA, global_batch: int = None) -> torch.Tensor:
    """pred_flows: [B,3,N_i] per scale (finest first); gt_flow [B,N,3]; fps_idxs int32 [B,N_{i+1}].
    CUDA fp32 inputs with a constant target: ONE fused kernel (forward + gradient, functional.kd_loss);
    otherwise (a target that needs a gradient) the reference op chain on kdpc gathers.
    ``global_batch`` (batch-sharded training): the mean over B (loss_functions.py:22) is taken over the GLOBAL batch, so
    that the SUM of the ranks' losses / gradients is the single-process loss / gradient over the concatenated batch."""
    B = gt_flow.shape[0]
    if _fusable(pred_flows, gt_flow, fps_idxs):
        return KF.kd_loss(pred_flows, fps_idxs, [gt_flow], [1.0 / (global_batch or B)], alpha)
    return multiScaleLoss_composed(pred_flows, gt_flow, fps_idxs, alpha) * (B / float(global_batch or B))


def multiScaleLoss_composed(pred_flows: Sequence[torch.Tensor], gt_flow: torch.Tensor, fps_idxs: Sequence[torch.Tensor],
                            alpha: Sequence[float] = ALPHA) -> torch.Tensor:
    """The reference's op chain (loss_functions.py:6-25) on the kdpc gather (differentiable w.r.t. everything)."""
    num_scale = len(pred_flows)
    offset = len(fps_idxs) - num_scale + 1
    gts: List[torch.Tensor] = [gt_flow]
    for idx in fps_idxs:
        gts.append(KF.gather_rows(gts[-1], idx) / scale)
    total = torch.zeros(1, device=gt_flow.device, dtype=gt_flow.dtype)
    for i in range(num_scale):
        diff = pred_flows[i].permute(0, 2, 1) - gts[i + offset]
        total = total + alpha[i] * torch.norm(diff, dim=2).sum(dim=1).mean()
    return total


def epe3d(pred_flow0: torch.Tensor, gt_flow: torch.Tensor) -> torch.Tensor:
    """distilTrain.py:229: mean over B*N of ||pred - gt||_2; pred_flow0 [B,3,N], gt [B,N,3]."""
    return torch.norm(pred_flow0.permute(0, 2, 1) - gt_flow, dim=2).mean()


def _kd_fused(outputs, fps_idxs, gt_flow, t0, w_teacher, w_gt, hints, alpha, global_batch=None):
    B = global_batch or gt_flow.shape[0]
    return KF.kd_loss(outputs, fps_idxs, [t0, gt_flow], [w_teacher / B, w_gt / B], alpha, hints)


def loss_fn_kd_2(outputs, fps_idxs, gt_flow, teacher_outputs, teacher_fps_idxs, gamma, alpha=ALPHA, global_batch=None):
    t0 = teacher_outputs[0].permute(0, 2, 1)
    if _fusable(outputs, gt_flow, fps_idxs) and not t0.requires_grad:
        return _kd_fused(outputs, fps_idxs, gt_flow, t0, gamma, 1 - gamma, (), alpha, global_batch)
    return (gamma * multiScaleLoss(outputs, t0, fps_idxs, alpha, global_batch)
            + (1 - gamma) * multiScaleLoss(outputs, gt_flow, fps_idxs, alpha, global_batch))


def biDirection_loss_ht(outputs, feat1s, feat2s, fps_idxs1, fps_idxs2, gt_flow, teacher_outputs, t_feat1s, t_feat2s,
                        t_fps_idxs1, t_fps_idxs2, gamma, beta, layer=0, alpha=ALPHA, global_batch=None):
    t0 = teacher_outputs[0].permute(0, 2, 1)
    if (_fusable(outputs, gt_flow, fps_idxs1, [(feat1s[layer], t_feat1s[layer]), (feat2s[layer], t_feat2s[layer])])
            and not t0.requires_grad):
        hints = [(feat1s[layer], t_feat1s[layer], 0.5 * (1 - beta)), (feat2s[layer], t_feat2s[layer], 0.5 * (1 - beta))]
        return _kd_fused(outputs, fps_idxs1, gt_flow, t0, beta * gamma, beta * (1 - gamma), hints, alpha, global_batch)
    loss1 = multiScaleLoss(outputs, t0, fps_idxs1, alpha, global_batch)
    loss2 = multiScaleLoss(outputs, gt_flow, fps_idxs1, alpha, global_batch)
    src = ((feat1s[layer] - t_feat1s[layer]) ** 2) / 2
    tgt = ((feat2s[layer] - t_feat2s[layer]) ** 2) / 2
    return beta * (gamma * loss1 + (1 - gamma) * loss2) + (1 - beta) * (0.5 * src.sum() + 0.5 * tgt.sum())


def cross_biDirection_loss_ht(outputs, feat1s, feat2s, fps_idxs1, fps_idxs2, gt_flow, teacher_outputs, t_feat1s,
                              t_feat2s, t_fps_idxs1, t_fps_idxs2, gamma, beta, layer=(2, 3), alpha=ALPHA,
                              hint_mode: str = "cat", global_batch=None):
    """``hint_mode='cat'`` is the reference formula verbatim (student feature vs cat(teacher feat1,
    teacher feat2) on channels — needs a student with twice the teacher's channels);
    ``hint_mode='first'`` compares against the teacher's feat1 only (shape-valid for the shipped
    student, same structure: MS-vs-teacher + MS-vs-GT + half squared hint error).
    NOTE the reference's hint term is a SUM over the batch while the flow terms are batch MEANS: under batch sharding
    the ranks' gradients must therefore be SUMMED with the means taken over ``global_batch`` (sharding.FlatGradAllReduce
    mode='sum') - averaging per-rank gradients would halve the hint term's share at 2 ranks."""
    t0 = teacher_outputs[0].permute(0, 2, 1)
    pairs = [(feat1s[e], torch.cat([t_feat1s[e], t_feat2s[e]], dim=1) if hint_mode == "cat" else t_feat1s[e]) for e in layer]
    if _fusable(outputs, gt_flow, fps_idxs1, pairs) and not t0.requires_grad:
        hints = [(a, b, 1 - beta) for a, b in pairs]
        return _kd_fused(outputs, fps_idxs1, gt_flow, t0, beta * gamma, beta * (1 - gamma), hints, alpha, global_batch)
    loss1 = multiScaleLoss(outputs, t0, fps_idxs1, alpha, global_batch)
    loss2 = multiScaleLoss(outputs, gt_flow, fps_idxs1, alpha, global_batch)
    hint = torch.zeros(1, device=gt_flow.device, dtype=gt_flow.dtype)
    for each in layer:
        t = torch.cat([t_feat1s[each], t_feat2s[each]], dim=1) if hint_mode == "cat" else t_feat1s[each]
        hint = hint + ((feat1s[each] - t) ** 2).sum() / 2
    return beta * (gamma * loss1 + (1 - gamma) * loss2) + (1 - beta) * hint
