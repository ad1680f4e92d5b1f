#!/usr/bin/env python
"""Diagnostic: whole-model gradient agreement of the unchanged student file on the kdpc kernels (compat stack) against
the stock reference stack on the same GPU, by parameter tensor, for eval-/train-mode BatchNorm and plain / KD losses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from oracle import ref_gpu
from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict

DEV = "cuda:0"
stock, compat = ref_gpu.load("stock"), ref_gpu.load("compat")
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = make_pairs(1, npts, seed=77, device=DEV)


def model(mods, name, seed):
    m = getattr(mods[name], "PointConvBidirection")()
    m.load_state_dict(synthetic_state_dict(m.state_dict(), seed))
    return m.to(DEV)


def grads(mods, train_bn, kd, tc_training=True):
    KF.clear_caches()
    KF.USE_TC_TRAINING = tc_training
    t, s = model(mods, "models_bid_pointconv", 7).eval(), model(mods, "models_bid_lighttoken_res", 8)
    s.train() if train_bn else s.eval()
    LF = mods["loss_functions"]
    with torch.no_grad():
        to = t(d["pos1"], d["pos2"], d["color1"], d["color2"])
    so = s(d["pos1"], d["pos2"], d["color1"], d["color2"])
    if kd:
        loss = LF.biDirection_loss_ht(so[0], so[5], so[6], so[1], so[2], d["flow"], to[0], to[5], to[6], to[1], to[2], 0.3, 0.8, layer=3)
    else:
        loss = LF.multiScaleLoss(so[0], d["flow"], so[1])
    loss.backward()
    KF.clear_caches()
    KF.USE_TC_TRAINING = True
    return loss.item(), {k: p.grad.detach().clone() for k, p in s.named_parameters() if p.grad is not None}, [f.detach() for f in so[0]]


for train_bn, kd, tc in ((False, False, True), (True, False, True), (True, True, True), (True, True, False)):
    l_ref, g_ref, f_ref = grads(stock, train_bn, kd)
    l_my, g_my, f_my = grads(compat, train_bn, kd, tc)
    l_ref2, g_ref2, _ = grads(stock, train_bn, kd)              # the stock stack against ITSELF (atomics / cuBLAS run-to-run)
    rows = sorted((((g_my[k] - g_ref[k]).norm() / g_ref[k].norm().clamp_min(1e-30)).item(),
                   ((g_ref2[k] - g_ref[k]).norm() / g_ref[k].norm().clamp_min(1e-30)).item(), k) for k in g_ref)
    fl = [((a - b).abs() > 1e-4 * b.abs().max()).float().mean().item() for a, b in zip(f_my, f_ref)]
    print(f"--- train_bn={train_bn} kd_loss={kd} tc_training={tc}: loss ref {l_ref:.6f} mine {l_my:.6f}  flow frac_bad {['%.4f' % x for x in fl]}")
    print(f"    rel-L2 gradient error by tensor: median {rows[len(rows) // 2][0]:.2e}  max {rows[-1][0]:.2e}   (stock vs stock: median "
          f"{sorted(r[1] for r in rows)[len(rows) // 2]:.2e} max {max(r[1] for r in rows):.2e})")
    for e, e2, k in rows[-8:]:
        print(f"      {e:.3e} (self {e2:.1e})  {k}")
