"""KD training step (training.py): the CUDA-graph replay is the eager step, loss for loss."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_graphed_kd_step_equals_eager_step():
    from kd_pointcloud_b200 import flownet, training
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
    torch.manual_seed(0)
    teacher = flownet.teacher()
    teacher.load_state_dict(synthetic_state_dict(teacher.state_dict(), 0))
    student = flownet.student()
    student.load_state_dict(synthetic_state_dict(student.state_dict(), 1))
    teacher, student = teacher.to(DEV), student.to(DEV)
    batches = [make_pairs(1, 2048, seed=70 + i, device=DEV) for i in range(3)]
    init = copy.deepcopy(student.state_dict())

    def run(graphed: bool):
        student.load_state_dict(init)
        opt = torch.optim.Adam(student.parameters(), lr=1e-4, capturable=True)
        losses = []
        if graphed:
            # the constructor's warm-up steps train too: rewind the student and the optimizer state afterwards
            stepper = training.GraphedKDStep(teacher, student, batches[0], opt)
            assert stepper.graph is not None
            student.load_state_dict(init)
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            step = stepper.step
        else:
            step = lambda b: training.kd_step(teacher, student, b, opt)
        for b in batches:
            losses.append(step(b).clone())
        return torch.cat([l.reshape(1) for l in losses]).cpu(), copy.deepcopy(student.state_dict())

    l_eager, sd_eager = run(False)
    l_graph, sd_graph = run(True)
    assert torch.isfinite(l_eager).all() and l_eager[0] > 0
    assert torch.equal(l_eager, l_graph), (l_eager, l_graph)
    for k in sd_eager:
        assert torch.equal(sd_eager[k], sd_graph[k]), k
