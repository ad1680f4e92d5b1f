"""Generate tests/golden/selfsup.npz with the UNMODIFIED reference functions (models_bid_pointconv.py:565-677, imported
from /root/reference on CPU exactly as tests/make_golden.py imports the layers) and assert that oracle/selfsup_ref.py
reproduces every one of them bit for bit.  Run in the build container only:  python tests/make_golden_selfsup.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import import_reference            # noqa: E402
from oracle import selfsup_ref as O                 # noqa: E402
from kd_pointcloud_b200.synth import make_pairs     # noqa: E402


def main():
    _, _, RM = import_reference()
    torch.manual_seed(0)
    d = make_pairs(2, 512, seed=21)
    g = torch.Generator().manual_seed(5)
    # a 3-level pyramid (512, 128, 64 points) with plausible flows: ground truth + noise
    pc1, pc2, flows = [], [], []
    for n in (512, 128, 64):
        sel = torch.randperm(512, generator=g)[:n]
        p1 = d["pos1"][:, sel].permute(0, 2, 1).contiguous()
        p2 = d["pos2"][:, torch.randperm(512, generator=g)[:n]].permute(0, 2, 1).contiguous()
        fl = (d["flow"][:, sel] + 0.05 * torch.randn(2, n, 3, generator=g)).permute(0, 2, 1).contiguous()
        pc1.append(p1); pc2.append(p2); flows.append(fl)
    out = {}
    with torch.no_grad():
        ref = RM.multiScaleChamferSmoothCurvature(pc1, pc2, flows)
        mine = O.multi_scale_chamfer_smooth_curvature(pc1, pc2, flows)
        for a, b, name in zip(ref, mine, ("total", "chamfer", "curvature", "smoothness")):
            assert torch.equal(a, b), (name, a, b)
            out[name] = a.numpy()
        parts = {"curvature_fn": (RM.curvature(pc2[0]), O.curvature(pc2[0])),
                 "smooth_fn": (RM.computeSmooth(pc1[0], flows[0]), O.compute_smooth(pc1[0], flows[0])),
                 "chamfer1_fn": (RM.computeChamfer(pc1[0] + flows[0], pc2[0])[0], O.compute_chamfer(pc1[0] + flows[0], pc2[0])[0]),
                 "chamfer2_fn": (RM.computeChamfer(pc1[0] + flows[0], pc2[0])[1], O.compute_chamfer(pc1[0] + flows[0], pc2[0])[1]),
                 "curvwarp_fn": (RM.curvatureWarp(pc1[0], pc1[0] + flows[0]), O.curvature_warp(pc1[0], pc1[0] + flows[0])),
                 "interp_fn": (RM.interpolateCurvature(pc1[0] + flows[0], pc2[0], RM.curvature(pc2[0])),
                               O.interpolate_curvature(pc1[0] + flows[0], pc2[0], O.curvature(pc2[0])))}
        for k, (a, b) in parts.items():
            assert torch.equal(a, b), k
            out[k] = a.numpy()
    for i in range(3):
        out[f"pc1_{i}"], out[f"pc2_{i}"], out[f"flow_{i}"] = pc1[i].numpy(), pc2[i].numpy(), flows[i].numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "selfsup.npz"), **out)
    print("wrote tests/golden/selfsup.npz", {k: v.shape for k, v in out.items() if not k.startswith(("pc", "flow_"))})


if __name__ == "__main__":
    main()
