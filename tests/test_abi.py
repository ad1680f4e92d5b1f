"""CPU tests: the C-ABI library loads and exports every symbol include/kdpc.h declares; the product
path has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kdpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kdpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kd_pointcloud_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libkdpc.so does not export {n}"
    assert sorted(_lib.exported_symbols()) == names        # the ctypes table covers the whole header
    assert L.kdpc_abi_version() == 1
    assert b"invalid" in L.kdpc_error_string(-1)


def test_null_arguments_are_rejected_without_a_gpu():
    from kd_pointcloud_b200 import _lib
    L = _lib.lib()
    assert L.kdpc_fps(1, 8, 4, None, None, None, None) == -1
    assert L.kdpc_knn(1, 8, 8, 64, None, None, None, None, None, None, None) == -1


def test_no_cpu_fallback():
    import kd_pointcloud_b200.pointnet2_utils as pu
    with pytest.raises(NotImplementedError):
        pu.furthest_point_sample(torch.zeros(1, 16, 3), 4)
    from kd_pointcloud_b200.pointconv_util import knn_point
    with pytest.raises(NotImplementedError):
        knn_point(3, torch.zeros(1, 16, 3), torch.zeros(1, 4, 3))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kd_pointcloud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() and "baseline/_ref" not in text, f"{f} mentions the oracle / reference install"


def test_compat_names_resolve_to_kdpc(monkeypatch):
    import sys
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    for m in ("pointconv_util", "pointconv_util2", "pointnet2", "pointnet2.pointnet2_utils", "thop", "pptk"):
        sys.modules.pop(m, None)
    import pointconv_util
    import pointconv_util2
    from pointnet2 import pointnet2_utils
    import thop  # noqa: F401
    import pptk  # noqa: F401
    for name in ("PointConv", "PointConvD", "PointWarping", "UpsampleFlow", "CrossLayerLight", "BottleNeck",
                 "SceneFlowEstimatorResidual", "index_points_gather", "index_points_group", "Conv1d",
                 "square_distance", "knn_point", "WeightNet", "group", "group_query", "Conv2d"):
        assert getattr(pointconv_util, name).__module__.startswith("kd_pointcloud_b200"), name
        assert hasattr(pointconv_util2, name)
    for name in ("furthest_point_sample", "gather_operation", "grouping_operation", "three_nn", "three_interpolate",
                 "ball_query", "QueryAndGroup", "GroupAll"):
        assert hasattr(pointnet2_utils, name)
    for m in ("pointconv_util", "pointconv_util2", "pointnet2", "pointnet2.pointnet2_utils", "thop", "pptk"):
        sys.modules.pop(m, None)
