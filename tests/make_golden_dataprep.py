"""Golden vectors for the GPU input pipeline, produced by the UNMODIFIED reference transforms
(/root/reference/transforms/transforms.py:137-316) under fixed numpy seeds (build container only):

    python tests/make_golden_dataprep.py        ->  tests/golden/dataprep.npz

For every case the script also re-draws the random numbers with kd_pointcloud_b200.dataprep's host-side ``draw_*``
methods under the SAME seed and asserts that oracle/dataprep_ref.py with those draws reproduces the reference output:
this pins both the oracle and the claim that the mirror classes consume numpy's stream exactly like the reference."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import dataprep_ref as OD  # noqa: E402


def raw_pair(n, seed, zmax=50.0):
    r = np.random.RandomState(seed)
    z = r.uniform(1.0, zmax, n).astype(np.float32)
    pc1 = np.stack([r.uniform(-0.457, 0.457, n).astype(np.float32) * z, r.uniform(-0.257, 0.257, n).astype(np.float32) * z, z], 1)
    flow = (r.randn(n, 3) * 0.5).astype(np.float32)
    return pc1.astype(np.float32), (pc1 + flow).astype(np.float32)


def main():
    sys.path.insert(0, REF)
    import transforms.transforms as RT                     # the reference's
    sys.path.pop(0)
    for m in [k for k in sys.modules if k == "transforms" or k.startswith("transforms.")]:
        sys.modules.pop(m)
    from kd_pointcloud_b200 import dataprep as KD           # only the host-side draw logic is used here (no GPU)

    dp = {"DEPTH_THRESHOLD": 35.0, "NO_CORR": True}
    dp_corr = {"DEPTH_THRESHOLD": 35.0, "NO_CORR": False}
    tog = {"degree_range": 0.1745329252, "shift_range": 1.0, "scale_low": 0.95, "scale_high": 1.05, "jitter_sigma": 0.01, "jitter_clip": 0.02}
    p2 = {"degree_range": 0.05, "shift_range": 0.3, "jitter_sigma": 0.01, "jitter_clip": 0.02}
    out = {}
    cases = [("pd", "ProcessData", dp, 3000, 1024, 5), ("pd_corr", "ProcessData", dp_corr, 3000, 1024, 6),
             ("pd_few", "ProcessData", dp, 900, 1024, 7),                 # fewer survivors than num_points: replace=True
             ("aug", "Augmentation", dp, 3000, 1024, 8), ("aug_corr", "Augmentation", dp_corr, 2500, 512, 9)]
    for name, kind, dpa, n, npts, seed in cases:
        pc1, pc2 = raw_pair(n, seed)
        if kind == "ProcessData":
            ref_t = RT.ProcessData(dpa, npts, allow_less_points=False)
            mine = KD.ProcessData(dpa, npts, False, device="cpu")
        else:
            ref_t = RT.Augmentation(tog, p2, dpa, npts)
            mine = KD.Augmentation(tog, p2, dpa, npts, device="cpu")
        np.random.seed(100 + seed)
        r1, r2, rsf = ref_t([pc1.copy(), pc2.copy()])
        # the same stream through the mirror's draw methods + the oracle
        np.random.seed(100 + seed)
        aff = mine.draw_affine(n)
        if aff is None:
            q1, q2, qsf, cnt = OD.process_data(pc1, pc2, dpa["DEPTH_THRESHOLD"], np.arange(1), np.arange(1))
        else:
            q1, q2, qsf, cnt = OD.augmentation(pc1, pc2, aff["affine"], aff["jitter1"], aff["jitter2"], dpa["DEPTH_THRESHOLD"],
                                               np.arange(1), np.arange(1))
        s1, s2 = mine.draw_selection(cnt)
        if aff is None:
            q1, q2, qsf, cnt = OD.process_data(pc1, pc2, dpa["DEPTH_THRESHOLD"], s1, s2)
        else:
            q1, q2, qsf, cnt = OD.augmentation(pc1, pc2, aff["affine"], aff["jitter1"], aff["jitter2"], dpa["DEPTH_THRESHOLD"], s1, s2)
        assert np.array_equal(q1, r1) and np.array_equal(q2, r2) and np.array_equal(qsf, rsf), f"{name}: oracle / draw order differs"
        print(f"  {name}: {cnt} of {n} survive the mask, {npts} sampled; oracle + mirror draws reproduce the reference bit for bit")
        out.update({f"{name}_pc1_raw": pc1, f"{name}_pc2_raw": pc2, f"{name}_pc1": r1, f"{name}_pc2": r2, f"{name}_sf": rsf,
                    f"{name}_sel1": s1, f"{name}_sel2": s2, f"{name}_count": np.int32(cnt), f"{name}_seed": np.int32(100 + seed)})
        if aff is not None:
            out[f"{name}_affine"] = aff["affine"]
            out[f"{name}_jitter1"] = aff["jitter1"]
            if aff["jitter2"] is not None:
                out[f"{name}_jitter2"] = aff["jitter2"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dataprep.npz"), **out)
    print("wrote tests/golden/dataprep.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "dataprep.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
