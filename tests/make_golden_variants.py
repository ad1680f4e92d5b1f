"""Golden vectors for the remaining layer variants, produced by the UNMODIFIED reference on CPU
(build container only):   python tests/make_golden_variants.py

  variants.npz   NoCrossLayerLight (pointconv_util.py:1276-1331), CrossLayerLightFG (:1871-1957) and PointConvWeight
                 (pointconv_util2.py:434-481) forwards with synthetic weights; oracle/layers_ref.py is asserted to
                 reproduce each of them while generating.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import close, import_reference, load_weights, save  # noqa: E402
from oracle import layers_ref as O  # noqa: E402
from kd_pointcloud_b200.synth import make_pairs  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    R, RL, RM = import_reference()
    import pointconv_util2 as R2                       # the reference's (sys.path has /root/reference first)
    d = make_pairs(2, 256, seed=13)
    pc1, pc2 = d["pos1"].permute(0, 2, 1).contiguous(), d["pos2"].permute(0, 2, 1).contiguous()
    g = torch.Generator().manual_seed(5)
    f1, f2 = torch.randn(2, 24, 256, generator=g) * 0.5, torch.randn(2, 24, 256, generator=g) * 0.5
    k1, k2 = torch.randn(2, 20, 256, generator=g), torch.randn(2, 20, 256, generator=g)
    out = {"pc1": pc1, "pc2": pc2, "feat1": f1, "feat2": f2, "knn1": k1, "knn2": k2}

    nc = R.NoCrossLayerLight(32, 24, [16, 16])
    sd = load_weights(nc, 11)
    y = nc(pc1, pc2, f1, f2)
    close(O.no_cross_layer_light({"n." + k: v for k, v in sd.items()}, "n", 32, pc1, pc2, f1, f2), y, 1e-5, "NoCrossLayerLight")
    out["nocross"] = y

    fg = R.CrossLayerLightFG(32, 24, [16, 16], [16, 16])
    sd = load_weights(fg, 12)
    a, b, c = fg(pc1, pc2, f1, f2, k1, k2)
    oa, ob, oc = O.cross_layer_light_fg({"c." + k: v for k, v in sd.items()}, "c", 32, pc1, pc2, f1, f2, k1, k2)
    for x, yy, nm in ((oa, a, "f1"), (ob, b, "f2"), (oc, c, "f3")):
        close(x, yy, 1e-5, "CrossLayerLightFG " + nm)
    out.update(fg1=a, fg2=b, fg3=c)
    # the feature-space neighbour sets themselves (16 nearest of knn1 in knn2), sorted by index
    idx = R.knn_point(16, k2.permute(0, 2, 1), k1.permute(0, 2, 1))
    mine = O.knn_point_feat(16, k2.permute(0, 2, 1), k1.permute(0, 2, 1))
    close(torch.sort(mine, dim=-1)[0], torch.sort(idx, dim=-1)[0], what="feature-space kNN")
    out["knn_feat16"] = torch.sort(idx, dim=-1)[0]

    xyz = pc1
    pts = torch.randn(2, 29, 256, generator=g) * 0.5
    pw = R2.PointConvWeight(64, 16, 29 + 3, 40)
    sd = load_weights(pw, 13)
    nx, ny, fidx = pw(xyz, pts)
    ox, oy, oi = O.pointconvd({"p." + k: v for k, v in sd.items()}, "p", 64, 16, xyz, pts)
    close(oi, fidx, what="PointConvWeight fps")
    close(oy, ny, 1e-5, "PointConvWeight feats")
    out.update(pcw_points=pts, pcw_new_xyz=nx, pcw_out=ny, pcw_fps=fidx)
    save("variants", **out)
    print("golden vectors written")


if __name__ == "__main__":
    main()
