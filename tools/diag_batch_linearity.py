#!/usr/bin/env python
"""Diagnostic: gradient of the KD loss over a batch of 4 == mean of the gradients over its two halves (BN in eval mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
from kd_pointcloud_b200 import functional as KF, flownet, losses
from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
dev = "cuda:0"
t, s = flownet.teacher(), flownet.student()
t.load_state_dict(synthetic_state_dict(t.state_dict(), 0)); s.load_state_dict(synthetic_state_dict(s.state_dict(), 1))
t, s = t.to(dev).eval(), s.to(dev).eval()
full = make_pairs(4, 2048, seed=31, device=dev)

def grads(batch):
    KF.clear_caches()
    s.zero_grad(set_to_none=True)
    with torch.no_grad():
        to = t(batch["pos1"], batch["pos2"], batch["color1"], batch["color2"])
    so = s(batch["pos1"], batch["pos2"], batch["color1"], batch["color2"])
    loss = losses.cross_biDirection_loss_ht(so[0], so[5], so[6], so[1], so[2], batch["flow"], to[0], to[5], to[6], to[1], to[2], 0.3, 0.8, layer=(2, 3), hint_mode="first")
    loss.backward()
    return loss.item(), {k: p.grad.clone() for k, p in s.named_parameters() if p.grad is not None}

for mode in sys.argv[1:] or ["default"]:
    KF.USE_TC_DW = "nodw" not in mode
    KF.USE_TC_TRAINING = "notc" not in mode
    KF.USE_FUSED_WEIGHTNET_GRAD = "nown" not in mode
    lf, gf = grads(full)
    halves = [grads({k: v[a:a + 2].contiguous() for k, v in full.items()}) for a in (0, 2)]
    rows = sorted((((0.5 * (halves[0][1][k] + halves[1][1][k]) - gf[k]).abs().max() / gf[k].abs().max().clamp_min(1e-30)).item(), k) for k in gf)
    print(f"--- {mode}: loss full {lf:.6f} halves {halves[0][0]:.6f} {halves[1][0]:.6f} mean {(halves[0][0] + halves[1][0]) / 2:.6f}")
    for e, k in rows[-10:]:
        print(f"   {e:.3e} {k}")
