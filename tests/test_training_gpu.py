"""KD training step (training.py): the CUDA-graph replay is the eager step, loss for loss; building the stepper does not
train; an eval forward after replays sees the trained weights; the captured learning rate can be rescheduled."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _models():
    from kd_pointcloud_b200 import flownet
    from kd_pointcloud_b200.synth import synthetic_state_dict
    torch.manual_seed(0)
    teacher = flownet.teacher()
    teacher.load_state_dict(synthetic_state_dict(teacher.state_dict(), 0))
    student = flownet.student()
    student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
    return teacher.to(DEV), student.to(DEV)


def test_graphed_kd_step_equals_eager_step():
    from kd_pointcloud_b200 import training
    from kd_pointcloud_b200.synth import make_pairs
    teacher, student = _models()
    batches = [make_pairs(1, 2048, seed=70 + i, device=DEV) for i in range(3)]
    init = copy.deepcopy(student.state_dict())

    def run(graphed: bool):
        student.load_state_dict(init)
        opt = torch.optim.Adam(student.parameters(), lr=1e-4, capturable=True)
        losses = []
        if graphed:
            stepper = training.GraphedKDStep(teacher, student, batches[0], opt)
            assert stepper.graph is not None
            # the constructor's warm-up steps are undone: weights, BN statistics and Adam state are where they started
            for k, v in student.state_dict().items():
                assert torch.equal(v, init[k]), k
            assert all(float(st["step"]) == 0 for st in opt.state.values())
            step = stepper.step
        else:
            step = lambda b: training.kd_step(teacher, student, b, opt)
        for b in batches:
            losses.append(step(b))                          # (step() returns a clone: no .clone() needed by the caller)
        return torch.cat([l.reshape(1) for l in losses]).cpu(), copy.deepcopy(student.state_dict())

    l_eager, sd_eager = run(False)
    l_graph, sd_graph = run(True)
    assert torch.isfinite(l_eager).all() and l_eager[0] > 0
    assert torch.equal(l_eager, l_graph), (l_eager, l_graph)
    for k in sd_eager:
        assert torch.equal(sd_eager[k], sd_graph[k]), k


def test_eval_after_graph_replays_uses_the_trained_weights():
    """replays move weights without moving tensor versions; the fused inference path must not hit packed
    weights / folded BN affines / WeightNet host parameters cached before (or between) the replays."""
    from kd_pointcloud_b200 import functional as KF
    from kd_pointcloud_b200 import training
    from kd_pointcloud_b200.synth import make_pairs
    teacher, student = _models()
    batches = [make_pairs(1, 2048, seed=80 + i, device=DEV) for i in range(3)]
    probe = make_pairs(1, 2048, seed=99, device=DEV)

    def eval_flow():
        student.eval()
        with torch.no_grad():
            KF.clear_caches()
            return student(probe["pos1"], probe["pos2"], probe["color1"], probe["color2"])[0][0].clone()

    opt = training.make_capturable_adam(student.parameters(), lr=1e-3)
    stepper = training.GraphedKDStep(teacher, student, batches[0], opt)
    assert stepper.graph is not None and len(stepper._keepalive) > 0
    before = eval_flow()                                    # fills every weight-derived cache with the initial weights
    for b in batches:
        stepper.step(b)
    after = eval_flow()
    KF.clear_caches(weights=True)                           # ground truth: everything re-derived from the weights
    truth = eval_flow()
    assert not torch.equal(before, truth)                   # training moved the weights
    assert torch.equal(after, truth)
    # the graph still replays correctly after the weight caches were dropped (it keeps its cached operands alive)
    l1 = stepper.step(batches[0])
    assert torch.isfinite(l1).all()

    # learning-rate schedule after capture (distilTrain.py:130-140): lr is a device tensor the captured Adam reads
    w0 = copy.deepcopy(student.state_dict())
    training.set_lr(opt, 0.0)
    stepper.step(batches[1])
    w1 = student.state_dict()
    assert all(torch.equal(w0[k], w1[k]) for k in w0 if "running" not in k and "tracked" not in k)   # lr = 0: no update
    training.set_lr(opt, 1e-3)
    stepper.step(batches[1])
    assert any(not torch.equal(w0[k], v) for k, v in student.state_dict().items() if k.endswith("weight"))

    # a short last batch (different shape) falls back to the eager step and the graph keeps working afterwards
    short = make_pairs(1, 2048, seed=5, device=DEV)
    short = {k: torch.cat([v, v], 0) for k, v in short.items()}
    assert torch.isfinite(stepper.step(short)).all()
    assert torch.isfinite(stepper.step(batches[2])).all()


def test_kdpc_adam_matches_torch_adam():
    """The one-pass Adam kernel (csrc/adam.cu) against torch.optim.Adam over five steps: ragged tensor sizes (chunk tails,
    unaligned element counts), weight decay, a learning-rate change between steps, more than one launch worth of tensors."""
    from kd_pointcloud_b200.training import KdpcAdam, set_lr
    g = torch.Generator().manual_seed(11)
    shapes = [(128, 2096), (33,), (7, 5, 3), (4096,), (1,), (257, 129)] + [(17 + i,) for i in range(120)]
    init = [torch.randn(s, generator=g) for s in shapes]
    for wd in (0.0, 1e-4):
        pa = [t.clone().to(DEV).requires_grad_(True) for t in init]
        pb = [t.clone().to(DEV).requires_grad_(True) for t in init]
        oa = KdpcAdam(pa, lr=1e-3, weight_decay=wd)
        ob = torch.optim.Adam(pb, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        for step in range(5):
            grads = [torch.randn(s, generator=g).to(DEV) * (1.0 + step) for s in shapes]
            for p, q, gr in zip(pa, pb, grads):
                p.grad, q.grad = gr.clone(), gr.clone()
            if step == 3:
                set_lr(oa, 5e-4)
                set_lr(ob, 5e-4)
            oa.step()
            ob.step()
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, rtol=2e-5, atol=1e-7), (p - q).abs().max()
