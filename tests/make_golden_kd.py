"""Golden vectors for the distillation losses (a19) and the BENCHMARK-shape whole model (a20), produced by the
UNMODIFIED reference on CPU (build container only):

    python tests/make_golden_kd.py

  kd_losses.npz          loss_fn_kd_2, biDirection_loss_ht, cross_biDirection_loss_ht of /root/reference/loss_functions.py
                         (:27-36, :83-96, :201-219): loss VALUES and autograd GRADIENTS w.r.t. the student's flows and
                         features.  cross_biDirection_loss_ht is evaluated verbatim (student hint features with twice
                         the teacher's channels, the only shapes for which the reference formula does not raise).
  model_teacher_n8192.npz  models_bid_pointconv.PointConvBidirection forward at the benchmark shape: B=1, N=8192,
                         pair 0 of bench.py's first batch (make_pairs(8, 8192, seed=1234)), MODEL_SEED weights.

While generating, oracle/layers_ref.py is asserted to reproduce the reference (pins the oracle at this shape too).
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from make_golden import close, import_reference, load_weights, save  # noqa: E402
from oracle import layers_ref as O  # noqa: E402
from kd_pointcloud_b200.synth import make_pairs  # noqa: E402


def kd_losses(RL):
    g = torch.Generator().manual_seed(123)
    B, sizes = 3, (1024, 256, 64, 32)
    preds = [torch.randn(B, 3, n, generator=g).requires_grad_(True) for n in sizes]
    fps = [torch.stack([torch.randperm(sizes[i], generator=g)[:sizes[i + 1]] for _ in range(B)]).int() for i in range(3)]
    gt = torch.randn(B, sizes[0], 3, generator=g)
    t_flow0 = torch.randn(B, 3, sizes[0], generator=g)
    chans = (8, 12, 16, 20)
    npts = (1024, 256, 64, 32)
    t1 = [torch.randn(B, c, n, generator=g) for c, n in zip(chans, npts)]
    t2 = [torch.randn(B, c, n, generator=g) for c, n in zip(chans, npts)]
    s1 = [torch.randn(B, c, n, generator=g).requires_grad_(True) for c, n in zip(chans, npts)]
    s2 = [torch.randn(B, c, n, generator=g).requires_grad_(True) for c, n in zip(chans, npts)]
    s1_wide = [torch.randn(B, 2 * c, n, generator=g).requires_grad_(True) for c, n in zip(chans, npts)]   # 'cat' form

    out = {"gt": gt, "t_flow0": t_flow0}
    for i in range(4):
        out.update({f"pred{i}": preds[i].detach(), f"t1_{i}": t1[i], f"t2_{i}": t2[i], f"s1_{i}": s1[i].detach(),
                    f"s2_{i}": s2[i].detach(), f"s1w_{i}": s1_wide[i].detach()})
    for i in range(3):
        out[f"fps{i}"] = fps[i]

    def run(name, fn, leaves):
        with torch.enable_grad():
            for t in leaves.values():
                t.grad = None
            loss = fn()
            loss.backward()
        out[f"{name}_loss"] = loss.detach()
        for k, t in leaves.items():
            out[f"{name}_g_{k}"] = torch.zeros_like(t) if t.grad is None else t.grad.clone()
        print(f"  {name}: {loss.item():.6f}")

    pl = {f"pred{i}": preds[i] for i in range(4)}
    run("kd2", lambda: RL.loss_fn_kd_2(preds, fps, gt, [t_flow0], None, 0.3), pl)
    run("bidir", lambda: RL.biDirection_loss_ht(preds, s1, s2, fps, fps, gt, [t_flow0], t1, t2, None, None, 0.3, 0.8, layer=1),
        {**pl, "s1_1": s1[1], "s2_1": s2[1]})
    run("cross", lambda: RL.cross_biDirection_loss_ht(preds, s1_wide, s2, fps, fps, gt, [t_flow0], t1, t2, None, None, 0.3, 0.8, layer=[2, 3]),
        {**pl, "s1w_2": s1_wide[2], "s1w_3": s1_wide[3]})
    # pin the oracle's restatements (oracle/layers_ref.py) on the same inputs
    with torch.no_grad():
        close(O.loss_fn_kd_2(preds, fps, gt, t_flow0, 0.3), out["kd2_loss"], 1e-6, "loss_fn_kd_2")
        close(O.bidirection_loss_ht(preds, s1, s2, fps, gt, t_flow0, t1, t2, 0.3, 0.8, layer=1), out["bidir_loss"], 1e-6,
              "biDirection_loss_ht")
        close(O.cross_bidirection_loss_ht(preds, s1_wide, fps, gt, t_flow0, t1, t2, 0.3, 0.8, layer=[2, 3]), out["cross_loss"],
              1e-6, "cross_biDirection_loss_ht")
    # the shipped shapes raise in the reference: record that fact
    try:
        RL.cross_biDirection_loss_ht(preds, s1, s2, fps, fps, gt, [t_flow0], t1, t2, None, None, 0.3, 0.8, layer=[2, 3])
        raised = 0
    except RuntimeError:
        raised = 1
    out["cross_equal_width_raises"] = torch.tensor(raised)
    assert raised == 1
    save("kd_losses", **out)


def model_8192(RL, RM):
    torch.set_grad_enabled(False)
    model = RM.PointConvBidirection()
    sd = load_weights(model, 7)                               # bench.py MODEL_SEED
    d8 = make_pairs(8, 8192, seed=1234)                       # bench.py's first batch on rank 0
    d = {k: v[:1].contiguous() for k, v in d8.items()}
    out = model(d["pos1"], d["pos2"], d["color1"], d["color2"])
    flows, fps1, fps2, pcs1, pcs2, feat1s, feat2s, crosses = out
    o = O.bid_pointconv_forward(sd, d["pos1"], d["pos2"], d["color1"], d["color2"], knn_impl="torch")
    for grp in (0, 5, 6, 7):
        for a, b in zip(o[grp], out[grp]):
            close(a, b, 0.0, f"8192-pt model output group {grp} (torch-kNN oracle)")
    o = O.bid_pointconv_forward(sd, d["pos1"], d["pos2"], d["color1"], d["color2"], knn_impl="c")
    for a, b in zip(o[0], flows):
        bad = ((a - b).abs() > 1e-4 * b.abs().max()).float().mean().item()
        assert bad < 5e-3, f"8192-pt flows: {bad:.2e} of elements differ"
    epe = torch.norm(flows[0].permute(0, 2, 1) - d["flow"], dim=2).mean()
    o_epe = torch.norm(o[0][0].permute(0, 2, 1) - d["flow"], dim=2).mean()
    assert abs(o_epe.item() - epe.item()) < 1e-4
    loss = RL.multiScaleLoss(flows, d["flow"], fps1)
    print(f"  8192-pt model: EPE3D {epe.item():.6f} (oracle with (distance,index) kNN: {o_epe.item():.6f}), loss {loss.item():.5f}")
    save("model_teacher_n8192", flow0=flows[0], flow1=flows[1], flow2=flows[2], flow3=flows[3],
         fps1_0=fps1[0], fps1_1=fps1[1], fps1_2=fps1[2], fps2_0=fps2[0], fps2_1=fps2[1], fps2_2=fps2[2],
         cross0_head=crosses[0][:, :, :512], cross3=crosses[3], loss=loss, epe3d=epe)


def main():
    torch.manual_seed(0)
    R, RL, RM = import_reference()
    kd_losses(RL)
    model_8192(RL, RM)
    print("golden vectors written")


if __name__ == "__main__":
    main()
