"""Evaluation metrics: oracle/eval_ref.py against the golden vectors made with the unmodified
reference (tests/make_golden_eval.py), and the CUDA kernel (csrc/metrics.cu) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import eval_ref as O


def test_oracle_matches_reference_golden(golden):
    g = golden("eval_metrics")
    for name in ("ft3d", "kitti"):
        mine = np.array(O.scene_flow_metrics(g[f"{name}_pc1"], g[f"{name}_pred"], g[f"{name}_gt"]))
        assert np.array_equal(mine, g[f"{name}_metrics"])
    mine = np.array(O.scene_flow_metrics(g["kitti_pc1"], g["kitti_pred"], g["kitti_gt"], g["kitti_calib"]))
    assert np.array_equal(mine, g["kitti_calib_metrics"])


def test_oracle_known_answers():
    gt = np.zeros((1, 4, 3), np.float32)
    gt[0, :, 0] = [1.0, 1.0, 1.0, 1.0]
    pred = gt.copy()
    pred[0, :, 0] += np.array([0.0, 0.04, 0.09, 0.5], np.float32)     # exact, strict, relaxed only, outlier
    epe, s, r, o = O.evaluate_3d(pred, gt)
    assert (s, r, o) == (0.5, 0.75, 0.25) and abs(epe - 0.1575) < 1e-6


def _check(m, ref, items):
    m = m.double().cpu().numpy()
    ref = np.asarray(ref, dtype=np.float64)
    assert abs(m[0] - ref[0]) <= 2e-6 * abs(ref[0]) + 1e-9 and abs(m[4] - ref[4]) <= 2e-6 * abs(ref[4]) + 1e-9   # error means
    for i in (1, 2, 3, 5):                                    # accuracy COUNTS are exact
        assert round(m[i] * items) == round(ref[i] * items), (i, m[i], ref[i])


@pytest.mark.gpu
def test_kernel_matches_oracle_on_golden(golden):
    from kd_pointcloud_b200 import evaluation_utils as E
    g = golden("eval_metrics")
    dev = "cuda:0"
    for name, calib in (("ft3d", None), ("kitti", None), ("kitti", "kitti_calib")):
        pc1, pred, gt = (torch.from_numpy(g[f"{name}_{k}"]).to(dev) for k in ("pc1", "pred", "gt"))
        cal = None if calib is None else torch.from_numpy(g[calib]).to(dev)
        ref = g[f"{name}_metrics" if calib is None else "kitti_calib_metrics"]
        items = pc1.shape[0] * pc1.shape[1]
        _check(E.scene_flow_metrics(pc1, pred, gt, cal), ref, items)                                   # point-major prediction
        m_cm = E.scene_flow_metrics(pc1, pred.permute(0, 2, 1).contiguous(), gt, cal)                  # the model's [B,3,N]
        _check(m_cm, ref, items)
        assert torch.equal(m_cm, E.scene_flow_metrics(pc1, pred.permute(0, 2, 1).contiguous(), gt, cal))   # deterministic
        e3 = E.evaluate_3d(g[f"{name}_pred"], g[f"{name}_gt"])                                         # reference-style numpy call
        _check(torch.tensor(list(e3) + [ref[4], ref[5]]), ref, items)
    fp, fg = O.get_batch_2d_flow(g["ft3d_pc1"], g["ft3d_pc1"] + g["ft3d_gt"], g["ft3d_pc1"] + g["ft3d_pred"])
    e2 = E.evaluate_2d(fp, fg)
    assert abs(e2[0] - g["ft3d_metrics"][4]) < 1e-5 * g["ft3d_metrics"][4] and abs(e2[1] - g["ft3d_metrics"][5]) < 1e-9


@pytest.mark.gpu
def test_kernel_full_size_and_meter():
    from kd_pointcloud_b200 import evaluation_utils as E
    from kd_pointcloud_b200.synth import make_pairs
    dev = "cuda:0"
    meter, refs = E.MetricMeter(), []
    for step in range(3):
        d = make_pairs(8, 8192, seed=40 + step)
        gen = torch.Generator().manual_seed(step)
        pred = d["flow"] + torch.randn(d["flow"].shape, generator=gen) * 0.1 * step
        ref = O.scene_flow_metrics(d["pos1"].numpy(), pred.numpy(), d["flow"].numpy())
        m = meter.update(d["pos1"].to(dev), pred.permute(0, 2, 1).contiguous().to(dev), d["flow"].to(dev))
        _check(m, ref, 8 * 8192)
        refs.append(ref)
    res = meter.result()
    mean = np.mean(np.array(refs), axis=0)
    for i, k in enumerate(E.NAMES):
        assert abs(res[k] - mean[i]) <= 1e-5 * abs(mean[i]) + 1e-7
    # exact prediction: zero error, every point accurate, no outliers
    m = E.scene_flow_metrics(d["pos1"].to(dev), d["flow"].to(dev), d["flow"].to(dev)).tolist()
    assert m[0] == 0.0 and m[1] == 1.0 and m[2] == 1.0 and m[3] == 0.0 and m[4] == 0.0 and m[5] == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(1, 1), (3, 1001), (2, 37), (5, 4099)])
def test_kernel_ragged_sizes(B, N):
    from kd_pointcloud_b200 import evaluation_utils as E
    g = torch.Generator().manual_seed(B * 100 + N)
    pc1 = torch.rand(B, N, 3, generator=g) * 20 + torch.tensor([0.0, 0.0, 5.0])
    gt = torch.randn(B, N, 3, generator=g) * 0.5
    pred = gt + torch.randn(B, N, 3, generator=g) * 0.1
    calib = torch.tensor([[-721.5, 609.6, 172.9, 44.9, 0.22, 0.0027]]).repeat(B, 1)
    for cal in (None, calib):
        ref = O.scene_flow_metrics(pc1.numpy(), pred.numpy(), gt.numpy(), None if cal is None else cal.numpy())
        m = E.scene_flow_metrics(pc1.cuda(), pred.permute(0, 2, 1).contiguous().cuda(), gt.cuda(), None if cal is None else cal.cuda())
        _check(m, ref, B * N)


@pytest.mark.gpu
def test_kernel_rejects_bad_shapes():
    from kd_pointcloud_b200 import evaluation_utils as E
    dev = "cuda:0"
    gt = torch.zeros(2, 16, 3, device=dev)
    with pytest.raises((ValueError, RuntimeError)):
        E.scene_flow_metrics(None, torch.zeros(2, 5, 16, device=dev), gt)
    with pytest.raises((ValueError, RuntimeError)):
        E.scene_flow_metrics(torch.zeros(2, 8, 3, device=dev), gt, gt)
    with pytest.raises(Exception):
        torch.ops.kdpc.flow_metrics(gt.cpu(), gt.cpu(), None, None, True)      # no CPU kernel registered
