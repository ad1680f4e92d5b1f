"""ORACLE (test infrastructure, never on the product path): numpy restatement of the reference's evaluation
metrics.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Follows evaluation_utils.py:17-50 (evaluate_3d, evaluate_2d) and utils/geometry.py:6-65 (get_batch_2d_flow,
project_3d_to_2d) of the reference; the only change is ``np.float`` -> ``np.float64`` (the alias was removed in
numpy 1.24, so the reference's own functions raise on this image's numpy 2.3 - SURVEY section 9).  Pinned against the
unmodified reference functions (run with the alias restored) by tests/make_golden_eval.py -> tests/golden/eval_metrics.npz.
"""
import numpy as np


def _fraction(mask) -> float:
    return mask.astype(np.float64).mean()


def _errors(pred, gt, eps):
    """per-point error norm and error relative to the ground-truth norm (+eps), float32 like the reference"""
    err = np.linalg.norm(gt - pred, axis=-1)
    return err, err / (np.linalg.norm(gt, axis=-1) + eps)


def evaluate_3d(sf_pred, sf_gt):
    """evaluation_utils.py:17-33.  (..., 3) float32 -> EPE3D, Acc3DS (5 cm or 5 %), Acc3DR (10 cm or 10 %),
    Outliers3D (> 30 cm or > 10 %)."""
    err, rel = _errors(sf_pred, sf_gt, 1e-4)
    return (err.mean(), _fraction((err < 0.05) | (rel < 0.05)), _fraction((err < 0.1) | (rel < 0.1)),
            _fraction((err > 0.3) | (rel > 0.1)))


def evaluate_2d(flow_pred, flow_gt):
    """evaluation_utils.py:36-50.  (..., 2) float32 -> EPE2D, Acc2D (3 px or 5 %)."""
    err, rel = _errors(flow_pred, flow_gt, 1e-5)
    return err.mean(), _fraction((err < 3.) | (rel < 0.05))


def project_3d_to_2d(pc, f=-1050., cx=479.5, cy=269.5, constx=0, consty=0, constz=0):
    """utils/geometry.py:61-65"""
    x = (pc[..., 0] * f + cx * pc[..., 2] + constx) / (pc[..., 2] + constz)
    y = (pc[..., 1] * f + cy * pc[..., 2] + consty) / (pc[..., 2] + constz)
    return x, y


def get_batch_2d_flow(pc1, pc2, predicted_pc2, calib=None):
    """utils/geometry.py:6-58 with the calibration already parsed: ``calib`` None (FlyingThings3D defaults) or a
    float32 [B,6] array (f = -P_rect[0,0], cx, cy, constx, consty, constz per sample, :25-38)."""
    if calib is not None:
        c = [np.asarray(calib[:, i], dtype=np.float32)[:, None] for i in range(6)]
        kw = dict(f=c[0], cx=c[1], cy=c[2], constx=c[3], consty=c[4], constz=c[5])
    else:
        kw = {}
    px1, py1 = project_3d_to_2d(pc1, **kw)
    px2, py2 = project_3d_to_2d(predicted_pc2, **kw)
    px2_gt, py2_gt = project_3d_to_2d(pc2, **kw)
    flow_x, flow_y = px2 - px1, py2 - py1
    flow_x_gt, flow_y_gt = px2_gt - px1, py2_gt - py1
    flow_pred = np.concatenate((flow_x[..., None], flow_y[..., None]), axis=-1)
    flow_gt = np.concatenate((flow_x_gt[..., None], flow_y_gt[..., None]), axis=-1)
    return flow_pred, flow_gt


def scene_flow_metrics(pc1, pred_sf, gt_sf, calib=None):
    """The six numbers evaluate_bid_pointconv.py:128-145 logs for one batch: pc1, pred_sf, gt_sf float32 [B,N,3]."""
    e3 = evaluate_3d(pred_sf, gt_sf)
    fp, fg = get_batch_2d_flow(pc1, pc1 + gt_sf, pc1 + pred_sf, calib)
    e2 = evaluate_2d(fp, fg)
    return tuple(float(v) for v in e3 + e2)
