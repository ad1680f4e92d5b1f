"""Generate tests/golden/eval_metrics.npz with the UNMODIFIED reference metric code (evaluation_utils.py,
utils/geometry.py under /root/reference), and assert that oracle/eval_ref.py reproduces it exactly.

Run in the build container only:  python tests/make_golden_eval.py
The reference uses ``np.float`` (removed in numpy 1.24); the alias is restored for this process only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
np.float = float                                     # noqa: NPY001  (what the reference was written against)

import evaluation_utils as RE                        # noqa: E402  the reference's
from utils import geometry as RG                     # noqa: E402
from oracle import eval_ref as O                     # noqa: E402
from kd_pointcloud_b200.synth import make_pairs      # noqa: E402


def main():
    out = {}
    for name, shape, seed in (("ft3d", "ft3d", 11), ("kitti", "kitti", 12)):
        d = make_pairs(3, 2048, seed=seed, kind=shape)
        pc1, gt = d["pos1"].numpy(), d["flow"].numpy()
        rng = np.random.default_rng(seed)
        # predictions at every error scale the thresholds separate (exact, 2 cm, 8 cm, 25 cm, 1 m noise)
        noise = rng.standard_normal(gt.shape).astype(np.float32) * rng.choice(
            np.array([0, 0.02, 0.08, 0.25, 1.0], dtype=np.float32), size=gt.shape[:2] + (1,))
        pred = (gt + noise).astype(np.float32)
        e3 = RE.evaluate_3d(pred, gt)
        # FlyingThings3D path of get_batch_2d_flow (paths without "KITTI": default intrinsics)
        fp, fg = RG.get_batch_2d_flow(pc1, pc1 + gt, pc1 + pred, ["/data/ft3d/0000"] * 3)
        e2 = RE.evaluate_2d(fp, fg)
        ref = np.array([float(v) for v in e3 + e2], dtype=np.float64)
        mine = np.array(O.scene_flow_metrics(pc1, pred, gt), dtype=np.float64)
        assert np.array_equal(ref, mine), (ref, mine)
        out.update({f"{name}_pc1": pc1, f"{name}_gt": gt, f"{name}_pred": pred, f"{name}_metrics": ref})
    # KITTI calibration path (geometry.py:7-38 reads P_rect_02 from calib_cam_to_cam/<name>.txt): same projection with
    # per-sample intrinsics; the reference's parsing is file I/O, its arithmetic is project_3d_to_2d with float32 arrays
    calib = np.array([[-721.5377, 609.5593, 172.854, 44.85728, 0.2163791, 0.002745884],
                      [-718.856, 607.1928, 185.2157, 45.38225, -0.1130887, 0.003779761],
                      [-707.0912, 601.8873, 183.1104, 46.88783, 0.1178601, 0.006203223]], dtype=np.float32)
    pc1, gt, pred = out["kitti_pc1"], out["kitti_gt"], out["kitti_pred"]
    kw = {k: calib[:, i][:, None] for i, k in enumerate(("f", "cx", "cy", "constx", "consty", "constz"))}
    px1, py1 = RG.project_3d_to_2d(pc1, **kw)
    px2, py2 = RG.project_3d_to_2d(pc1 + pred, **kw)
    pxg, pyg = RG.project_3d_to_2d(pc1 + gt, **kw)
    fp = np.stack([px2 - px1, py2 - py1], -1)
    fg = np.stack([pxg - px1, pyg - py1], -1)
    e2 = RE.evaluate_2d(fp, fg)
    ref = np.array([float(v) for v in RE.evaluate_3d(pred, gt) + e2], dtype=np.float64)
    mine = np.array(O.scene_flow_metrics(pc1, pred, gt, calib), dtype=np.float64)
    assert np.array_equal(ref, mine), (ref, mine)
    out.update({"kitti_calib": calib, "kitti_calib_metrics": ref})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_metrics.npz"), **out)
    print("wrote tests/golden/eval_metrics.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
