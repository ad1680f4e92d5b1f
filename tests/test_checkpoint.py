"""Reference checkpoint layouts (kd_pointcloud_b200/checkpoint.py): plain state_dict, DataParallel prefixes, wrapped
dicts; validation errors name what differs.  CPU only."""
import json
import os

import pytest
import torch

from kd_pointcloud_b200 import checkpoint as C
from kd_pointcloud_b200 import flownet
from kd_pointcloud_b200.synth import synthetic_state_dict


@pytest.fixture(scope="module")
def model_and_sd():
    m = flownet.student()
    sd = synthetic_state_dict(m.state_dict(), 3)
    return m, sd


def test_reference_key_set_loads_unchanged(model_and_sd):
    m, sd = model_and_sd
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")))
    names = set(keys if isinstance(keys, list) else keys.keys())
    assert names == set(sd.keys())                       # the reference model's own parameter / buffer names
    rep = C.load_reference_checkpoint(m, sd)
    assert rep.ok and rep.matched == len(sd) and not rep.stripped_prefix
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])


def test_dataparallel_and_wrapped_layouts(model_and_sd, tmp_path):
    m, sd = model_and_sd
    dp = {"module." + k: v for k, v in sd.items()}
    rep = C.load_reference_checkpoint(m, dp)
    assert rep.ok and rep.stripped_prefix == "module."
    rep = C.load_reference_checkpoint(m, {"epoch": 7, "state_dict": {"module.module." + k: v for k, v in sd.items()}})
    assert rep.ok and rep.stripped_prefix == "module.module." and rep.unwrapped_key == "state_dict"
    path = tmp_path / "ckpt.pth"
    torch.save(dp, path)
    assert C.load_reference_checkpoint(m, str(path)).ok


def test_mismatches_are_reported_by_name(model_and_sd):
    m, sd = model_and_sd
    bad = dict(sd)
    gone = "level0.composed_module.0.weight"
    del bad[gone]
    bad["extra.weight"] = torch.zeros(3)
    bad["flow0.fc.weight"] = torch.zeros(3, 32, 1)
    nanned = "cross1.pos1.bias"
    bad[nanned] = sd[nanned].clone()
    bad[nanned][0] = float("nan")
    rep, _ = C.check_reference_checkpoint(m, bad)
    assert not rep.ok and rep.missing == [gone] and rep.unexpected == ["extra.weight"]
    assert [n for n, _, _ in rep.shape_mismatch] == ["flow0.fc.weight"] and rep.non_finite == [nanned]
    with pytest.raises(RuntimeError, match="flow0.fc.weight"):
        C.load_reference_checkpoint(m, bad)
    before = m.state_dict()["flow0.fc.weight"].clone()
    rep = C.load_reference_checkpoint(m, bad, strict=False)          # loads what fits, leaves the rest
    assert torch.equal(m.state_dict()["flow0.fc.weight"], before) and rep.matched == len(sd) - 3
    with pytest.raises(ValueError):
        C.check_reference_checkpoint(m, {"epoch": 1})
    with pytest.raises(TypeError):
        C.check_reference_checkpoint(m, [1, 2, 3])
