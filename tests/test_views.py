"""Host logic of the concat-free inference forward (CPU): which views the strided kernels accept (``ops.row_stride``),
when two tensors already are the halves of one (``functional.joined``), and that the switch only opens on the
no-grad CUDA path (``functional.concat_free``).  The kernels themselves: tests/test_tc_gpu.py, tests/test_ops_gpu.py."""
import torch

from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200 import ops


def test_row_stride_accepts_column_blocks_and_batch_halves_only():
    b = torch.empty(16, 100, 96)
    assert ops.row_stride(b) == 96
    assert ops.row_stride(b[:, :, :64]) == 96 and ops.row_stride(b[:, :, 64:]) == 96      # column blocks
    assert ops.row_stride(b[:8]) == 96 and ops.row_stride(b[8:, :, 32:64]) == 96           # batch halves (+ block)
    assert ops.row_stride(b[0:1, :, :32]) == 96                                            # size-1 axes are free
    assert ops.row_stride(b[:, :50, :64]) is None                                          # rows not uniformly strided
    assert ops.row_stride(b.permute(0, 2, 1)) is None                                      # channels not unit-strided
    assert ops.row_stride(torch.empty(5, 3).t()) is None
    assert ops.row_stride(torch.empty(7)) == 7 and ops.row_stride(torch.empty(0, 5, 8)) == 8


def test_joined_returns_the_parent_only_for_its_exact_halves():
    ab = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(6, 4, 5).clone()      # (an allocation of its own, not a view)
    a, b = ab[:3], ab[3:]
    assert KF.joined(a, b) is ab
    for x, y in ((b, a), (ab[:2], ab[2:4]), (a, b.clone()), (a.clone(), b), (ab[:3, :, :4], ab[3:, :, :4])):
        j = KF.joined(x, y)
        assert j is not ab and torch.equal(j, torch.cat([x, y], dim=0))
    other = torch.zeros(6, 4, 5)
    assert torch.equal(KF.joined(a, other[3:]), torch.cat([a, other[3:]], dim=0))


def test_concat_free_needs_cuda_float32_and_no_autograd():
    x = torch.zeros(2, 3)
    assert not KF.concat_free(x)                                    # CPU tensors never
    with torch.no_grad():
        assert not KF.concat_free(x)
        assert not KF.concat_free(torch.zeros(2, 3, device="meta"))  # not CUDA
