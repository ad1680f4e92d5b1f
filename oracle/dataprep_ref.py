"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's ProcessData / Augmentation
(transforms/transforms.py:137-194, :197-316) with the random draws passed in explicitly, so that parity is defined:
``sel1`` / ``sel2`` are positions in the survivor list (np.random.choice(indices, n, replace=False) ==
indices[np.random.permutation(len(indices))[:n]]).  Pinned against the unmodified reference classes run under the same
numpy seed by tests/make_golden_dataprep.py."""
from __future__ import annotations

import numpy as np


def _select(pc1, pc2, sf, depth_threshold, sel1, sel2):
    if depth_threshold > 0:
        near = np.logical_and(pc1[:, 2] < depth_threshold, pc2[:, 2] < depth_threshold)      # transforms.py:151-152
    else:
        near = np.ones(pc1.shape[0], dtype=bool)
    indices = np.where(near)[0]
    return pc1[indices[sel1]], pc2[indices[sel2]], sf[indices[sel1]], len(indices)


def process_data(pc1, pc2, depth_threshold, sel1, sel2):
    """ProcessData.__call__, transforms.py:144-192."""
    pc1, pc2 = pc1[:, :3].astype(np.float32), pc2[:, :3].astype(np.float32)
    sf = pc2 - pc1
    return _select(pc1, pc2, sf, depth_threshold, sel1, sel2)


def augmentation(pc1, pc2, affine, jitter1, jitter2, depth_threshold, sel1, sel2):
    """Augmentation.__call__, transforms.py:206-315.  affine = matrix(9) shifts(3) matrix2^T(9) shifts2(3) as built at
    :229-276; jitter1 the clipped noise of :250-253 (or None), jitter2 that of :281-285 (None when NO_CORR)."""
    pc1, pc2 = pc1[:, :3].astype(np.float32).copy(), pc2[:, :3].astype(np.float32).copy()
    a = np.asarray(affine, dtype=np.float32)
    matrix, shifts, matrix2_t, shifts2 = a[:9].reshape(3, 3), a[9:12].reshape(1, 3), a[12:21].reshape(3, 3), a[21:24].reshape(1, 3)
    bias = shifts + (0 if jitter1 is None else jitter1)
    pc1 = (pc1.dot(matrix) + bias).astype(np.float32)
    pc2 = (pc2.dot(matrix) + bias).astype(np.float32)
    pc2 = (pc2.dot(matrix2_t) + shifts2).astype(np.float32)
    sf = pc2 - pc1
    if jitter2 is not None:
        pc2 = pc2 + jitter2
    return _select(pc1, pc2, sf, depth_threshold, sel1, sel2)
