"""Edge cases of the drop-in API on the GPU: empty and degenerate inputs, the reference's own argument checks
(pointnet2_utils.py:22,50-51,89-90,121-123,167-168 ``assert is_contiguous``; torch.topk's k > N error behind knn_point,
pointconv_util.py:106), maximum sizes of each kNN / FPS path, and the no-CPU-fallback guarantee."""
import numpy as np
import pytest
import torch

from oracle import layers_ref as O
from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200 import pointconv_util as P
from kd_pointcloud_b200 import pointnet2_utils as PU
from kd_pointcloud_b200.synth import make_pairs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = torch.ops.kdpc


def test_empty_inputs_return_empty_outputs():
    xyz = make_pairs(2, 64, seed=1, device=DEV)["pos1"]
    assert PU.furthest_point_sample(xyz, 0).shape == (2, 0)                       # sampling_gpu.cu:100 (m <= 0: nothing)
    e = torch.zeros(0, 64, 3, device=DEV)
    assert PU.furthest_point_sample(e, 4).shape == (0, 4)
    assert KF.knn_idx(3, xyz, torch.zeros(2, 0, 3, device=DEV)).shape == (2, 0, 3)
    assert KF.gather_rows(torch.randn(2, 64, 8, device=DEV), torch.zeros(2, 0, dtype=torch.int32, device=DEV)).shape == (2, 0, 8)
    f = torch.randn(2, 5, 64, device=DEV)
    assert PU.gather_operation(f, torch.zeros(2, 0, dtype=torch.int32, device=DEV)).shape == (2, 5, 0)
    assert PU.grouping_operation(f, torch.zeros(2, 0, 4, dtype=torch.int32, device=DEV)).shape == (2, 5, 0, 4)


def test_reference_argument_checks_are_kept():
    xyz = make_pairs(2, 128, seed=2, device=DEV)["pos1"]
    with pytest.raises(AssertionError):                                           # pointnet2_utils.py:22
        PU.furthest_point_sample(xyz.permute(0, 2, 1).contiguous().permute(0, 2, 1), 16)
    with pytest.raises(RuntimeError):                                             # torch.topk: k out of range (pointconv_util.py:106)
        P.knn_point(200, xyz, xyz)
    with pytest.raises((TypeError, RuntimeError)):
        K.knn(xyz.double(), xyz.double(), 3)
    with pytest.raises((NotImplementedError, RuntimeError)):                      # no CPU kernels behind the ops
        K.fps(xyz.cpu(), 4)
    f = torch.randn(2, 4, 128, device=DEV)
    with pytest.raises(AssertionError):                                           # pointnet2_utils.py:167-168
        PU.grouping_operation(f, torch.zeros(2, 8, 4, dtype=torch.int32, device=DEV)[:, :, ::2])


def test_degenerate_clouds():
    one = torch.tensor([[[1.0, 2.0, 3.0]]], device=DEV)                           # a single point
    assert PU.furthest_point_sample(one, 1).tolist() == [[0]]
    assert KF.knn_idx(1, one, one).tolist() == [[[0]]]
    same = torch.ones(2, 300, 3, device=DEV)                                      # all points identical: ties everywhere
    fps = PU.furthest_point_sample(same, 7)
    assert torch.equal(fps.cpu(), O.furthest_point_sample(same.cpu(), 7))
    idx = KF.knn_idx(5, same, same[:, :10].contiguous())
    assert torch.equal(idx.cpu().long(), O.knn_point(5, same.cpu(), same[:, :10].cpu()))
    assert (idx == torch.arange(5, device=DEV)).all()                             # (distance, index): the 5 lowest indices
    dist, i3 = PU.three_nn(same[:, :4].contiguous(), same)
    assert (dist == 0).all() and i3.tolist()[0][0] == [0, 1, 2]


@pytest.mark.parametrize("n,s,k", [(16384, 4096, 32), (20000, 512, 16), (255, 255, 9), (256, 31, 3), (33, 33, 32)])
def test_knn_size_limits_of_every_path(n, s, k):
    """n = 16384: the largest pruned search; n > 16384 and n < 256: the TMA-tiled brute-force kernel; k = n edge."""
    g = torch.Generator().manual_seed(n)
    cand = (torch.rand(1, n, 3, generator=g) * 30).to(DEV)
    q = (torch.rand(1, s, 3, generator=g) * 30).to(DEV)
    KF.clear_caches()
    idx = KF.knn_idx(k, cand, q)
    ref = O.knn_point(k, cand.cpu(), q.cpu()[:, :64].contiguous())
    assert torch.equal(idx[:, :64].cpu().long(), ref)
    assert torch.equal(idx, K.knn_bruteforce(q, cand, k))
    KF.clear_caches()


def test_fps_largest_register_path_and_generic_path():
    for n, m in ((16384, 1024), (20000, 64), (31, 8)):
        xyz = make_pairs(1, n, seed=n, device=DEV)["pos1"]
        assert torch.equal(PU.furthest_point_sample(xyz, m).cpu(), O.furthest_point_sample(xyz.cpu(), m)), n
