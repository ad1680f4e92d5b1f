"""GPU input pipeline: ProcessData / Augmentation (transforms/transforms.py:137-316).
CPU: the numpy oracle against the golden the unmodified reference produced.  GPU: csrc/dataprep.cu through the mirror
classes, with the stored draws and with numpy's own stream under the stored seed."""
import numpy as np
import pytest
import torch

from oracle import dataprep_ref as OD

DP = {"DEPTH_THRESHOLD": 35.0, "NO_CORR": True}
DP_CORR = {"DEPTH_THRESHOLD": 35.0, "NO_CORR": False}
TOG = {"degree_range": 0.1745329252, "shift_range": 1.0, "scale_low": 0.95, "scale_high": 1.05, "jitter_sigma": 0.01, "jitter_clip": 0.02}
P2 = {"degree_range": 0.05, "shift_range": 0.3, "jitter_sigma": 0.01, "jitter_clip": 0.02}
CASES = [("pd", False, DP, 1024), ("pd_corr", False, DP_CORR, 1024), ("pd_few", False, DP, 1024), ("aug", True, DP, 1024),
         ("aug_corr", True, DP_CORR, 512)]


def _draws(g, name, aug):
    d = {"sel1": g[f"{name}_sel1"], "sel2": g[f"{name}_sel2"]}
    if aug:
        d.update(affine=g[f"{name}_affine"], jitter1=g[f"{name}_jitter1"], jitter2=g.get(f"{name}_jitter2"))
    return d


@pytest.mark.parametrize("name,aug,dp,npts", CASES)
def test_oracle_matches_reference_golden(golden, name, aug, dp, npts):
    g = golden("dataprep")
    d = _draws(g, name, aug)
    if aug:
        p1, p2, sf, cnt = OD.augmentation(g[f"{name}_pc1_raw"], g[f"{name}_pc2_raw"], d["affine"], d["jitter1"], d["jitter2"],
                                          dp["DEPTH_THRESHOLD"], d["sel1"], d["sel2"])
    else:
        p1, p2, sf, cnt = OD.process_data(g[f"{name}_pc1_raw"], g[f"{name}_pc2_raw"], dp["DEPTH_THRESHOLD"], d["sel1"], d["sel2"])
    assert cnt == int(g[f"{name}_count"]) and p1.shape == (npts, 3)
    assert np.array_equal(p1, g[f"{name}_pc1"]) and np.array_equal(p2, g[f"{name}_pc2"]) and np.array_equal(sf, g[f"{name}_sf"])


def _make(aug, dp, npts):
    from kd_pointcloud_b200 import dataprep as KD
    return KD.Augmentation(TOG, P2, dp, npts, device="cuda:0") if aug else KD.ProcessData(dp, npts, False, device="cuda:0")


def _check(out, g, name, aug):
    p1, p2, sf = (t.cpu().numpy() for t in out)
    if aug:     # 3x3 products: numpy's float32 dot may contract differently (1 ulp); indices and structure are exact
        for a, b in ((p1, g[f"{name}_pc1"]), (p2, g[f"{name}_pc2"])):
            assert np.allclose(a, b, rtol=2e-6, atol=2e-6)
        assert np.allclose(sf, g[f"{name}_sf"], rtol=0, atol=1e-5)
    else:
        assert np.array_equal(p1, g[f"{name}_pc1"]) and np.array_equal(p2, g[f"{name}_pc2"]) and np.array_equal(sf, g[f"{name}_sf"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,aug,dp,npts", CASES)
def test_gpu_pipeline_matches_reference_golden(golden, name, aug, dp, npts):
    g = golden("dataprep")
    t = _make(aug, dp, npts)
    raw = (g[f"{name}_pc1_raw"], g[f"{name}_pc2_raw"])
    _check(t(raw, draws=_draws(g, name, aug)), g, name, aug)          # explicit draws
    assert int(t.last_counts[0]) == int(g[f"{name}_count"])
    np.random.seed(int(g[f"{name}_seed"]))                            # numpy's own stream, the reference's call order
    _check(t(raw), g, name, aug)


@pytest.mark.gpu
def test_gpu_pipeline_batches_ragged_samples_and_rejects_bad_draws(golden):
    g = golden("dataprep")
    t = _make(False, DP, 1024)
    pairs = [(g["pd_pc1_raw"], g["pd_pc2_raw"]), (g["pd_few_pc1_raw"], g["pd_few_pc2_raw"])]       # 3000 and 900 raw points
    out = t.batch(pairs, draws=[_draws(g, "pd", False), _draws(g, "pd_few", False)])
    assert out["pos1"].shape == (2, 1024, 3) and torch.equal(out["color1"], out["pos1"])
    for i, name in enumerate(("pd", "pd_few")):
        _check((out["pos1"][i], out["pos2"][i], out["flow"][i]), g, name, False)
    assert list(t.last_counts) == [int(g["pd_count"]), int(g["pd_few_count"])]
    bad = _draws(g, "pd", False)
    bad["sel1"] = bad["sel1"].copy()
    bad["sel1"][3] = int(g["pd_count"])                               # one past the survivor list
    with pytest.raises(ValueError):
        t.batch(pairs[:1], draws=[bad])
    far = (g["pd_pc1_raw"] + np.float32(100.0), g["pd_pc2_raw"] + np.float32(100.0))               # nothing survives the mask
    assert t(far) == (None, None, None)
