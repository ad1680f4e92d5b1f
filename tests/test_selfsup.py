"""Self-supervised loss family: oracle/selfsup_ref.py against the golden vectors of the unmodified
reference (tests/make_golden_selfsup.py), and kd_pointcloud_b200/selfsup.py (kdpc kNN + gather kernels) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import selfsup_ref as O

T = torch.from_numpy


def _pyramid(g):
    return ([T(g[f"pc1_{i}"]) for i in range(3)], [T(g[f"pc2_{i}"]) for i in range(3)], [T(g[f"flow_{i}"]) for i in range(3)])


def test_oracle_matches_reference_golden(golden):
    g = golden("selfsup")
    pc1, pc2, flows = _pyramid(g)
    out = O.multi_scale_chamfer_smooth_curvature(pc1, pc2, flows)
    for v, name in zip(out, ("total", "chamfer", "curvature", "smoothness")):
        assert np.array_equal(v.numpy(), g[name])
    assert np.array_equal(O.curvature(pc2[0]).numpy(), g["curvature_fn"])
    assert np.array_equal(O.compute_smooth(pc1[0], flows[0]).numpy(), g["smooth_fn"])
    assert np.array_equal(O.interpolate_curvature(pc1[0] + flows[0], pc2[0], O.curvature(pc2[0])).numpy(), g["interp_fn"])


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.gpu
def test_gpu_functions_match_oracle(golden):
    from kd_pointcloud_b200 import selfsup as S
    g = golden("selfsup")
    dev = "cuda:0"
    pc1, pc2, flows = _pyramid(g)
    p1, p2, fl = pc1[0].to(dev), pc2[0].to(dev), flows[0].to(dev)
    # neighbour sets are exact, so everything that only SUMS neighbours agrees to fp32 rounding
    assert _rel(S.curvature(p2), g["curvature_fn"]) < 1e-5
    assert _rel(S.curvatureWarp(p1, p1 + fl), g["curvwarp_fn"]) < 1e-5
    assert _rel(S.computeSmooth(p1, fl), g["smooth_fn"]) < 1e-5
    # distances by value: the matmul expansion carries ~6e-5 of absolute rounding noise at these coordinates (|p|^2 ~ 1e3)
    # and the reference adds the two norms in (pc1, pc2) order whichever cloud is the query - dist2 differs by that noise
    d1, d2 = S.computeChamfer(p1 + fl, p2)
    assert (d1.cpu() - T(g["chamfer1_fn"])).abs().max() < 2e-4 and (d2.cpu() - T(g["chamfer2_fn"])).abs().max() < 2e-4
    assert _rel(S.interpolateCurvature(p1 + fl, p2, S.curvature(p2)), g["interp_fn"]) < 2e-3
    out = S.multiScaleChamferSmoothCurvature([t.to(dev) for t in pc1], [t.to(dev) for t in pc2], [t.to(dev) for t in flows])
    for v, name, tol in zip(out, ("total", "chamfer", "curvature", "smoothness"), (1e-4, 1e-4, 1e-3, 1e-5)):
        assert abs(v.item() - float(g[name][0])) <= tol * abs(float(g[name][0])), (name, v.item(), g[name])


@pytest.mark.gpu
def test_gpu_gradients_match_oracle_autograd(golden):
    from kd_pointcloud_b200 import selfsup as S
    g = golden("selfsup")
    dev = "cuda:0"
    pc1, pc2, flows = _pyramid(g)
    fg = [f.clone().to(dev).requires_grad_(True) for f in flows]
    S.multiScaleChamferSmoothCurvature([t.to(dev) for t in pc1], [t.to(dev) for t in pc2], fg)[0].backward()
    fr = [f.clone().requires_grad_(True) for f in flows]
    O.multi_scale_chamfer_smooth_curvature(pc1, pc2, fr)[0].backward()
    for a, b in zip(fg, fr):
        # (a nearest-neighbour switch between the two evaluations would show up as an O(1) difference in one row)
        bad = ((a.grad.cpu() - b.grad).abs() > 1e-3 * b.grad.abs().max()).float().mean().item()
        assert bad < 5e-3, bad
