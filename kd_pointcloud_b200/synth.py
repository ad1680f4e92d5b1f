"""Synthetic scene-flow pairs shaped like the reference's datasets.

There is no dataset on the benchmark box; shapes and value ranges follow what the reference's
loaders produce (datasets/kitti.py:47-49,97-102, transforms/transforms.py:137-194,
utils/geometry.py:61): 8192 points per cloud after the depth < 35 m mask, NO_CORR resampling (no
index correspondence between the clouds), colour == xyz.  Generated on the CPU generator so that
the same seed gives the same pair on every machine, then moved to the requested device.
"""
from __future__ import annotations

from typing import Dict

import torch


def make_pairs(batch: int, npoints: int = 8192, seed: int = 1234, kind: str = "ft3d", device="cpu",
               duplicates: float = 0.0, quantize: float = 0.0) -> Dict[str, torch.Tensor]:
    """Returns dict(pos1, pos2, color1, color2, flow), each [B, npoints, 3] float32.

    kind='ft3d' : pin-hole frustum (f=1050, 960x540), depth U(1,35); rigid per-cluster motion.
    kind='kitti': x U(-15,21), y U(-1.4,2.4), z U(6,35); ego-motion dominated flow.
    duplicates  : fraction of points replaced by exact copies of others (replace=True resampling,
                  transforms.py:178) — the stress case for the FPS / kNN tie rules.
    quantize    : if > 0, coordinates are rounded to this grid (exact distance ties).
    """
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    B, N = batch, npoints
    if kind == "ft3d":
        z = torch.rand(B, N, generator=g) * 34.0 + 1.0
        u = (torch.rand(B, N, generator=g) * 2 - 1) * 0.457
        v = (torch.rand(B, N, generator=g) * 2 - 1) * 0.257
        pos1 = torch.stack([u * z, v * z, z], dim=2)
        # 16 clusters by x-quantile, each with its own rigid shift
        rank = pos1[..., 0].argsort(dim=1).argsort(dim=1)
        cluster = (rank * 16 // N).clamp(max=15)
        shifts = torch.randn(B, 16, 3, generator=g) * 0.5
        flow = torch.gather(shifts, 1, cluster.unsqueeze(-1).expand(B, N, 3)) + torch.randn(B, N, 3, generator=g) * 0.02
    elif kind == "kitti":
        lo = torch.tensor([-15.0, -1.4, 6.0])
        hi = torch.tensor([21.0, 2.4, 35.0])
        pos1 = torch.rand(B, N, 3, generator=g) * (hi - lo) + lo
        mean = torch.tensor([-0.01, 0.0, -1.7])
        std = torch.tensor([0.04, 0.01, 0.53])
        flow = torch.randn(B, N, 3, generator=g) * std + mean
    else:
        raise ValueError(kind)
    if quantize > 0:
        pos1 = torch.round(pos1 / quantize) * quantize
    if duplicates > 0:
        nd = int(N * duplicates)
        src = torch.randint(0, N, (B, nd), generator=g)
        dst = torch.randint(0, N, (B, nd), generator=g)
        pos1.scatter_(1, dst.unsqueeze(-1).expand(B, nd, 3), torch.gather(pos1, 1, src.unsqueeze(-1).expand(B, nd, 3)))
        flow.scatter_(1, dst.unsqueeze(-1).expand(B, nd, 3), torch.gather(flow, 1, src.unsqueeze(-1).expand(B, nd, 3)))
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    pos2 = torch.gather(pos1 + flow, 1, perm.unsqueeze(-1).expand(B, N, 3)).contiguous()
    pos1 = pos1.contiguous().float()
    out = {"pos1": pos1, "pos2": pos2.float(), "color1": pos1.clone(), "color2": pos2.float().clone(),
           "flow": flow.contiguous().float()}
    return {k: v.to(device) for k, v in out.items()}


def synthetic_state_dict(reference_state: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights that do not depend on module construction order: every tensor is
    drawn from a generator seeded by crc32(key), scaled like torch's default init.  Used so that
    the reference model (golden generation) and this package load the SAME weights without
    shipping a 32 MB checkpoint."""
    import math
    import zlib

    out = {}
    for key, ref in reference_state.items():
        g = torch.Generator(device="cpu").manual_seed((zlib.crc32(key.encode()) + seed) & 0x7fffffff)
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shape, dtype=ref.dtype)
        elif key.endswith("running_var"):
            out[key] = 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif key.endswith("running_mean"):
            out[key] = 0.05 * torch.randn(shape, generator=g)
        elif ref.dim() >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound          # torch's default U(-1/sqrt(fan_in), ..)
        else:
            # biases / BN affine
            if "bn" in key and key.endswith("weight"):
                out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                out[key] = 0.05 * torch.randn(shape, generator=g)
    return out
