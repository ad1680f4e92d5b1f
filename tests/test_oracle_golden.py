"""CPU tests: the oracle (oracle/) against the golden vectors produced by the unmodified
reference (tests/make_golden.py), and the oracle's internal consistency."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import layers_ref as O
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict

T = torch.from_numpy


def _rel(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_square_distance_and_knn_match_reference(golden):
    g = golden("knn")
    q, c = T(g["query"]), T(g["cand"])
    # bit-exact: the scalar-C evaluation order equals what torch's matmul expansion produced
    assert np.array_equal(O.square_distance_c(q, c).numpy(), g["sqdist"])
    for k in (3, 9, 16, 32):
        mine = torch.sort(O.knn_point(k, c, q), dim=-1)[0].numpy()
        assert np.array_equal(mine, g[f"knn{k}"])


def test_knn_oracle_is_sorted_by_distance_then_index():
    g = torch.Generator().manual_seed(3)
    cloud = torch.randint(0, 6, (1, 300, 3), generator=g).float()   # integer grid: many exact ties
    idx, dist = O.knn_with_dist(16, cloud, cloud[:, :50].contiguous())
    dd, ii = dist.numpy(), idx.numpy()
    assert np.all(np.diff(dd, axis=-1) >= 0)
    tie = np.diff(dd, axis=-1) == 0
    assert tie.any() and np.all(np.diff(ii, axis=-1)[tie] > 0)


@pytest.mark.parametrize("n,m", [(256, 64), (512, 256), (1000, 100), (2048, 512), (37, 9)])
def test_fps_literal_simulation_equals_closed_form_tie_rule(n, m):
    # integer-grid clouds are full of exact distance ties; duplicates too
    g = torch.Generator().manual_seed(n)
    xyz = torch.randint(0, 4, (2, n, 3), generator=g).float()
    a = O.furthest_point_sample(xyz, m)
    b = O.furthest_point_sample(xyz, m, closed_form=True)
    assert torch.equal(a, b)
    assert (a[:, 0] == 0).all()


def test_fps_tie_rule_is_not_lowest_index():
    # the reference's left-biased tree picks bit-reversed thread order, not the lowest index
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 1:, 0] = 1.0                                 # points 1..7 all at distance 1 from point 0
    idx = O.furthest_point_sample(xyz, 2)
    assert idx[0, 1].item() == 4                        # bitrev3: 4 -> 001 is the smallest non-zero key


def test_three_nn_oracle_properties():
    d = make_pairs(2, 128, seed=5)
    dist, idx = O.three_nn(d["pos1"], d["pos2"])
    assert dist.shape == (2, 128, 3) and idx.dtype == torch.int32
    assert (dist[..., 0] <= dist[..., 1]).all() and (dist[..., 1] <= dist[..., 2]).all()
    brute = torch.cdist(d["pos1"], d["pos2"]).topk(3, largest=False)[0]
    assert torch.allclose(dist, brute, atol=1e-3)      # cdist uses the matmul expansion (less exact)


def _sd(module_keys, seed, prefix):
    ref = {k: torch.zeros(v) for k, v in module_keys.items()}
    return {prefix + k: v for k, v in synthetic_state_dict(ref, seed).items()}


def _shapes(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def test_layers_match_reference_golden(golden):
    from kd_pointcloud_b200 import pointconv_util as P   # only used for parameter shapes (CPU, no kernels)

    g = golden("weightnet")
    sd = _sd(_shapes(P.WeightNet(3, 16)), 1, "w.")
    assert _rel(O.weightnet(sd, "w", T(g["localized_xyz"])), g["out"]) < 1e-6

    g = golden("pointconv")
    sd = _sd(_shapes(P.PointConv(9, 32, 24, bn=True)), 2, "p.")
    assert _rel(O.pointconv(sd, "p", 9, T(g["xyz"]), T(g["points"]), bn=True), g["out"]) < 1e-5

    g = golden("pointconvd")
    sd = _sd(_shapes(P.PointConvD(64, 16, 32, 40)), 3, "p.")
    nx, ny, fi = O.pointconvd(sd, "p", 64, 16, T(g["xyz"]), T(g["points"]))
    assert np.array_equal(fi.numpy(), g["fps_idx"]) and np.array_equal(nx.numpy(), g["new_xyz"])
    assert _rel(ny, g["out"]) < 1e-5

    g = golden("crosslayer")
    sd = _sd(_shapes(P.CrossLayerLight(32, 24, [16, 16], [16, 16])), 4, "c.")
    outs = O.cross_layer_light(sd, "c", 32, T(g["pc1"]), T(g["pc2"]), T(g["feat1"]), T(g["feat2"]))
    for o, name in zip(outs, ("out1", "out2", "out3")):
        assert _rel(o, g[name]) < 1e-5

    g = golden("warp_upsample")
    assert _rel(O.point_warping(T(g["pc1"]), T(g["pc2"]), T(g["flow1"])), g["warped"]) < 1e-6
    assert _rel(O.upsample_flow(T(g["pc1"]), T(g["sparse_xyz"]), T(g["sparse_flow"])), g["up"]) < 1e-6

    g = golden("flow_estimator")
    sd = _sd(_shapes(P.SceneFlowEstimatorResidual(24, 16, channels=[32, 32], mlp=[32, 16])), 5, "e.")
    f, fl = O.scene_flow_estimator_residual(sd, "e", T(g["xyz"]), T(g["feats"]), T(g["cost"]), T(g["flow"]))
    assert _rel(f, g["out_feat"]) < 1e-5 and _rel(fl, g["out_flow"]) < 1e-5

    g = golden("multiscale_loss")
    l = O.multi_scale_loss([T(g["p0"]), T(g["p1"]), T(g["p2"])], T(g["gt"]), [T(g["fps1"]), T(g["fps2"])])
    assert _rel(l, g["loss"]) < 1e-6


def test_state_dict_keys_match_reference():
    from kd_pointcloud_b200.flownet import PointConvBidirection
    with open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")) as f:
        ref = json.load(f)
    mine = {k: list(v.shape) for k, v in PointConvBidirection().state_dict().items()}
    assert list(mine) == list(ref)
    assert mine == ref
    assert sum(int(np.prod(v)) for k, v in mine.items() if "running" not in k and "tracked" not in k) == 7958604


def test_kd_loss_oracles_match_reference_golden(golden):
    """loss_fn_kd_2 / biDirection_loss_ht / cross_biDirection_loss_ht restatements against values AND gradients the
    unmodified loss_functions.py produced (tests/make_golden_kd.py)."""
    g = golden("kd_losses")
    fps = [T(g[f"fps{i}"]) for i in range(3)]
    gt, t0 = T(g["gt"]), T(g["t_flow0"])
    t1, t2 = [T(g[f"t1_{i}"]) for i in range(4)], [T(g[f"t2_{i}"]) for i in range(4)]

    def leaves(prefix):
        return [T(g[f"{prefix}{i}"]).clone().requires_grad_(True) for i in range(4)]

    preds, s1, s2, s1w = leaves("pred"), leaves("s1_"), leaves("s2_"), leaves("s1w_")
    cases = [("kd2", lambda: O.loss_fn_kd_2(preds, fps, gt, t0, 0.3), {}),
             ("bidir", lambda: O.bidirection_loss_ht(preds, s1, s2, fps, gt, t0, t1, t2, 0.3, 0.8, layer=1), {"s1_1": s1[1], "s2_1": s2[1]}),
             ("cross", lambda: O.cross_bidirection_loss_ht(preds, s1w, fps, gt, t0, t1, t2, 0.3, 0.8, layer=[2, 3]),
              {"s1w_2": s1w[2], "s1w_3": s1w[3]})]
    for name, fn, extra in cases:
        for t in preds + s1 + s2 + s1w:
            t.grad = None
        loss = fn()
        loss.backward()
        assert _rel(loss.detach(), g[f"{name}_loss"]) < 1e-6
        for i in range(4):
            assert _rel(preds[i].grad, g[f"{name}_g_pred{i}"]) < 1e-5
        for k, t in extra.items():
            assert _rel(t.grad, g[f"{name}_g_{k}"]) < 1e-6
    with pytest.raises(RuntimeError):                      # equal student/teacher widths: the reference raises
        O.cross_bidirection_loss_ht(preds, s1, fps, gt, t0, t1, t2, 0.3, 0.8, layer=[2, 3])
    assert int(g["cross_equal_width_raises"]) == 1


def test_layer_variant_oracles_match_reference_golden(golden):
    """NoCrossLayerLight / CrossLayerLightFG / PointConvWeight restatements against the unmodified
    reference's outputs (tests/make_golden_variants.py)."""
    from kd_pointcloud_b200 import pointconv_util as P
    g = golden("variants")
    pc1, pc2, f1, f2, k1, k2 = (T(g[k]) for k in ("pc1", "pc2", "feat1", "feat2", "knn1", "knn2"))
    sd = _sd(_shapes(P.NoCrossLayerLight(32, 24, [16, 16])), 11, "n.")
    assert _rel(O.no_cross_layer_light(sd, "n", 32, pc1, pc2, f1, f2), g["nocross"]) < 1e-5
    sd = _sd(_shapes(P.CrossLayerLightFG(32, 24, [16, 16], [16, 16])), 12, "c.")
    outs = O.cross_layer_light_fg(sd, "c", 32, pc1, pc2, f1, f2, k1, k2)
    for o, name in zip(outs, ("fg1", "fg2", "fg3")):
        assert _rel(o, g[name]) < 1e-5
    idx = torch.sort(O.knn_point_feat(16, k2.permute(0, 2, 1), k1.permute(0, 2, 1)), dim=-1)[0]
    assert np.array_equal(idx.numpy(), g["knn_feat16"])
    sd = _sd(_shapes(P.PointConvWeight(64, 16, 32, 40)), 13, "p.")
    nx, ny, fi = O.pointconvd(sd, "p", 64, 16, pc1, T(g["pcw_points"]))
    assert np.array_equal(fi.numpy(), g["pcw_fps"]) and _rel(ny, g["pcw_out"]) < 1e-5
