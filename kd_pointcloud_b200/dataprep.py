"""GPU input pipeline: ``ProcessData`` and ``Augmentation`` of the reference's loaders
(transforms/transforms.py:137-194 and :197-316) on batches, feeding ``runner.FlowRunner`` / ``training.kd_step``.

Same constructor arguments and the same random stream as the reference: the classes draw from ``np.random`` in the
reference's order (scale, angle, shifts, jitter, angle2, shifts2, [jitter2], choice, [choice]), so with the same numpy
seed a sample comes out as the reference produces it (coordinates to float32 rounding of the 3x3 products, indices
exactly).  The per-point work — two affine maps, flow, depth mask, np.where compaction, the fancy-index gathers — runs
in csrc/dataprep.cu; randomness stays on the host (a few scalars and one permutation per sample), or is supplied as
explicit ``draws`` (dicts with the keys of ``draw()``), which is how parity is defined.

``__call__((pc1, pc2))`` accepts numpy arrays or tensors [n, >=3] like the reference and returns DEVICE tensors
``(pc1 [num_points,3], pc2, sf)``; ``batch(list_of_pairs)`` processes a whole batch in three launches and returns
``dict(pos1, pos2, color1, color2, flow)`` in the layout the models take (colour = xyz, datasets/kitti.py:47-48).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import ops

_stream = ops._stream
_p = ops._p


def _as_f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a[:, :3], dtype=np.float32)


class _Prep:
    augment = False

    def __init__(self, data_process_args, num_points, allow_less_points=False, device="cuda"):
        self.DEPTH_THRESHOLD = float(data_process_args['DEPTH_THRESHOLD'])
        self.no_corr = bool(data_process_args['NO_CORR'])
        self.num_points = int(num_points)
        self.allow_less_points = allow_less_points
        self.device = torch.device(device)

    # -- randomness (host, numpy's global stream, the reference's call order) -----------------------------------------
    def draw_affine(self, n_raw: int) -> Optional[dict]:
        return None

    def draw_selection(self, count: int) -> Tuple[np.ndarray, np.ndarray]:
        """transforms.py:158-192 / :295-315: np.random.choice(indices, num_points, replace=False) twice when NO_CORR,
        with replacement when fewer than num_points survive (and allow_less_points is off).  Returned as POSITIONS in
        the survivor list: choice(indices, n, False) == indices[permutation(len)[:n]], (.., True) == indices[randint]."""
        n = self.num_points

        def one():
            if count >= n:
                return np.random.permutation(count)[:n]
            return np.random.randint(0, count, size=n)
        s1 = one()
        s2 = one() if self.no_corr else s1
        return s1.astype(np.int32), s2.astype(np.int32)

    # -- device work --------------------------------------------------------------------------------------------------
    def batch(self, pairs: Sequence[Tuple], draws: Optional[List[dict]] = None) -> Dict[str, torch.Tensor]:
        """pairs: [(pc1_raw [n_i, >=3], pc2_raw [n_i, >=3]), ...] (numpy or tensors).  draws[i] (optional): dict with
        'affine' (24 floats), 'jitter1', 'jitter2' ([n_i,3] or None), 'sel1', 'sel2' (int positions)."""
        if self.num_points <= 0:
            raise NotImplementedError("num_points <= 0 (variable-size outputs) is not supported on the batched GPU path")
        B = len(pairs)
        dev = self.device
        L = _lib.lib()
        raw = [(_as_f32(a), _as_f32(b)) for a, b in pairs]
        n_raw = np.array([a.shape[0] for a, _ in raw], dtype=np.int32)
        nmax = int(n_raw.max())
        h1 = torch.zeros(B, nmax, 3, pin_memory=True)
        h2 = torch.zeros(B, nmax, 3, pin_memory=True)
        for i, (a, b) in enumerate(raw):
            if a.shape != b.shape:
                raise ValueError("pc1 and pc2 must have the same number of points (the reference subtracts them row by row)")
            h1[i, :a.shape[0]] = torch.from_numpy(a)
            h2[i, :b.shape[0]] = torch.from_numpy(b)
        d1, d2 = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)
        if draws is None:
            affs = [self.draw_affine(int(n)) for n in n_raw]
        else:
            affs = [dr if self.augment else None for dr in draws]
        aff_t = j1_t = j2_t = None
        if self.augment:
            aff_t = torch.tensor(np.stack([np.asarray(a['affine'], dtype=np.float32) for a in affs]), device=dev)
            if affs[0].get('jitter1') is not None:
                j = np.zeros((B, nmax, 3), dtype=np.float32)
                for i, a in enumerate(affs):
                    j[i, :n_raw[i]] = a['jitter1']
                j1_t = torch.from_numpy(j).to(dev)
            if affs[0].get('jitter2') is not None:
                j = np.zeros((B, nmax, 3), dtype=np.float32)
                for i, a in enumerate(affs):
                    j[i, :n_raw[i]] = a['jitter2']
                j2_t = torch.from_numpy(j).to(dev)
        n_t = torch.from_numpy(n_raw).to(dev)
        ws = torch.empty(L.kdpc_dataprep_workspace_bytes(B, nmax), dtype=torch.uint8, device=dev)
        count = torch.empty(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ops._call("kdpc_dataprep_mask", B, nmax, 3, self.DEPTH_THRESHOLD, 1 if self.augment else 0, _p(n_t), _p(d1), _p(d2),
                      _p(aff_t), _p(j1_t), _p(j2_t), _p(ws), _p(count), _stream())
        counts = count.cpu().numpy()                        # one small device->host read per BATCH (numpy needs len(indices))
        if (counts == 0).any():
            raise ValueError("indices = np.where(mask)[0], len(indices) == 0")       # the reference prints this and returns None
        if draws is None:
            sels = [self.draw_selection(int(c)) for c in counts]
        else:
            sels = [(np.asarray(dr['sel1'], dtype=np.int32), np.asarray(dr['sel2'], dtype=np.int32)) for dr in draws]
        s1 = torch.from_numpy(np.stack([s[0] for s in sels])).to(dev)
        s2 = torch.from_numpy(np.stack([s[1] for s in sels])).to(dev)
        out = [torch.empty(B, self.num_points, 3, device=dev) for _ in range(3)]
        status = torch.zeros(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ops._call("kdpc_dataprep_select", B, nmax, self.num_points, _p(ws), _p(count), _p(s1), _p(s2), _p(out[0]), _p(out[1]),
                      _p(out[2]), _p(status), _stream())
        if draws is not None and bool(status.any()):
            raise ValueError("a supplied draw lies outside the survivor list")
        self.last_counts = counts
        return {"pos1": out[0], "pos2": out[1], "color1": out[0].clone(), "color2": out[1].clone(), "flow": out[2]}

    def __call__(self, data, draws: Optional[dict] = None):
        pc1, pc2 = data
        if pc1 is None:
            return None, None, None
        try:
            o = self.batch([(pc1, pc2)], None if draws is None else [draws])
        except ValueError as e:
            if "len(indices) == 0" in str(e):
                print('indices = np.where(mask)[0], len(indices) == 0')
                return None, None, None
            raise
        return o["pos1"][0], o["pos2"][0], o["flow"][0]


class ProcessData(_Prep):
    """transforms.py:137-194: depth mask + random subsample; flow = pc2 - pc1 of the raw clouds."""

    def __repr__(self):
        return (f"{self.__class__.__name__}\n(data_process_args: \n\tDEPTH_THRESHOLD: {self.DEPTH_THRESHOLD}\n\tNO_CORR: {self.no_corr}\n"
                f"\tallow_less_points: {self.allow_less_points}\n\tnum_points: {self.num_points}\n)")


class Augmentation(_Prep):
    """transforms.py:197-316: scale . rotation(y) + shift + jitter on both clouds, a second rotation + shift on cloud 2,
    flow recomputed, then ProcessData's mask and subsample."""
    augment = True

    def __init__(self, aug_together_args, aug_pc2_args, data_process_args, num_points, allow_less_points=False, device="cuda"):
        super().__init__(data_process_args, num_points, allow_less_points, device)
        self.together_args = aug_together_args
        self.pc2_args = aug_pc2_args

    def draw_affine(self, n_raw: int) -> dict:
        t, p = self.together_args, self.pc2_args
        scale = np.diag(np.random.uniform(t['scale_low'], t['scale_high'], 3).astype(np.float32))
        angle = np.random.uniform(-t['degree_range'], t['degree_range'])
        c, s = np.cos(angle), np.sin(angle)
        rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
        matrix = scale.dot(rot.T)
        shifts = np.random.uniform(-t['shift_range'], t['shift_range'], (1, 3)).astype(np.float32)
        jitter = np.clip(t['jitter_sigma'] * np.random.randn(n_raw, 3), -t['jitter_clip'], t['jitter_clip']).astype(np.float32)
        angle2 = np.random.uniform(-p['degree_range'], p['degree_range'])
        c2, s2 = np.cos(angle2), np.sin(angle2)
        matrix2 = np.array([[c2, 0, s2], [0, 1, 0], [-s2, 0, c2]], dtype=np.float32)
        shifts2 = np.random.uniform(-p['shift_range'], p['shift_range'], (1, 3)).astype(np.float32)
        jitter2 = None
        if not self.no_corr:
            jitter2 = np.clip(p['jitter_sigma'] * np.random.randn(n_raw, 3), -p['jitter_clip'], p['jitter_clip']).astype(np.float32)
        affine = np.concatenate([matrix.reshape(-1), shifts.reshape(-1), matrix2.T.reshape(-1), shifts2.reshape(-1)]).astype(np.float32)
        return {"affine": affine, "jitter1": jitter, "jitter2": jitter2}
