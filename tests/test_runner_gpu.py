"""CUDA-graph runner (kd_pointcloud_b200/runner.py): graph replay == eager forward, the end-to-end host calls, and the
software-pipelined streaming call give the same per-batch results; the metrics come from the fused kernel."""
import pytest
import torch

from oracle import eval_ref as OE

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pinned(d):
    return {k: v.contiguous().pin_memory() for k, v in d.items()}


def test_runner_graph_host_and_pipelined_calls_agree():
    from kd_pointcloud_b200 import flownet
    from kd_pointcloud_b200.runner import FlowRunner
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
    torch.manual_seed(0)
    model = flownet.teacher()
    model.load_state_dict(synthetic_state_dict(model.state_dict(), 0))
    model = model.to(DEV).eval()
    B, N = 2, 2048
    batches = [_pinned(make_pairs(B, N, seed=100 + i)) for i in range(5)]
    eager = FlowRunner(model, B, N, DEV, use_graph=False)
    eager.warmup_and_capture(batches[0], warmup=1)
    graph = FlowRunner(model, B, N, DEV, use_graph=True)
    assert graph.warmup_and_capture(batches[0], warmup=1)
    ref = [eager.run_host(b) for b in batches]
    flows = []
    for b, r in zip(batches, ref):
        assert graph.run_host(b) == r                              # bit-identical replay
        flows.append(graph.out_flow.clone())
    assert graph.run_host_pipelined(iter(batches)) == ref           # any number of batches, two staging sets
    assert graph.run_host_pipelined(iter(batches[:1])) == ref[:1]
    assert graph.run_host_pipelined(iter([])) == []
    # the six metrics of the last batch against the numpy oracle on the runner's own flow
    graph.run_host(batches[-1])
    m = graph.out_metrics.cpu().double().numpy()
    o = OE.scene_flow_metrics(batches[-1]["pos1"].numpy(), graph.out_flow.permute(0, 2, 1).contiguous().cpu().numpy(),
                              batches[-1]["flow"].numpy())
    assert abs(m[0] - o[0]) < 2e-6 * abs(o[0]) + 1e-9
    for i in (1, 2, 3, 5):
        assert round(m[i] * B * N) == round(o[i] * B * N)


@pytest.mark.parametrize("dual_forward", [False, True])
def test_pipelined_runner_equals_plain_runner(dual_forward):
    """Sampling pyramid of batch i+1 on a second stream beside the forward of batch i (one-CTA FPS kernel, persistent
    kernels capped to the SMs it leaves free): same metrics and flows as the single-graph runner, batch for batch."""
    from kd_pointcloud_b200 import flownet
    from kd_pointcloud_b200.runner import FlowRunner, PipelinedFlowRunner
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
    torch.manual_seed(0)
    model = flownet.teacher()
    model.load_state_dict(synthetic_state_dict(model.state_dict(), 0))
    model = model.to(DEV).eval()
    B, N = 2, 4096
    host = [_pinned(make_pairs(B, N, seed=300 + i)) for i in range(5)]
    dev_batches = [{k: v.to(DEV) for k, v in b.items()} for b in host]
    plain = FlowRunner(model, B, N, DEV, use_graph=True)
    assert plain.warmup_and_capture(host[0], warmup=1)
    ref, ref_flow = [], []
    for b in host:
        ref.append(plain.run_host(b))
        ref_flow.append(plain.out_flow.clone())
    ref_m = []
    for b in host:
        plain.run_host(b)
        ref_m.append(plain.out_metrics.clone())
    from kd_pointcloud_b200 import ops
    pipe = PipelinedFlowRunner(model, B, N, DEV, dual_forward=dual_forward)   # True: the two slots' forwards on their own streams
    assert pipe.warmup_and_capture(host[0], warmup=1) and 0 < pipe.sm_limit < 148
    assert pipe.precompute_knn and pipe.neighbours[0] is not None and len(pipe.neighbours[0][0]) == 14   # kNN sets moved to stream B
    for _ in range(2):                                                  # twice: slots and events are reused
        got = pipe.run_resident(dev_batches)
        torch.cuda.synchronize()
        for a, b in zip(got, ref_m):
            assert torch.equal(a, b)
    assert torch.equal(pipe.out_flow[(len(host) - 1) & 1], ref_flow[-1])
    assert pipe.run_host_pipelined(iter(host)) == ref
    assert pipe.run_host_pipelined(iter(host[:1])) == ref[:1]
    assert pipe.run_host_pipelined(iter([])) == []
    from kd_pointcloud_b200 import _lib
    assert _lib.lib().kdpc_sm_limit() == 0                              # the cap only applies while capturing
