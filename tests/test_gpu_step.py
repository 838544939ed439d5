"""GPU tests of the whole Graph-HSCN step: CUDA step vs the CPU oracle step, padded vs compact virtual
layout, CUDA-graph replay vs eager execution."""
import pytest
import torch

from tests.util import RTOL, assert_close, rel_err

pytestmark = pytest.mark.gpu


def _cfg():
    from graph_hscn_b200.train import StepConfig
    return StepConfig(hidden=48, num_layers=3, num_clusters=10, lr=1e-2)


def _sync_weights(dst_step, src_step):
    dst_step.scn.load_state_dict(src_step.scn.state_dict())
    dst_step.hscn.load_state_dict(src_step.hscn.state_dict())


def test_step_matches_oracle_step(cuda):
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.train import GraphHSCNStep
    from oracle.step import OracleStep
    batch = synthetic.peptides_batch(12, seed=77)
    ostep = OracleStep(_cfg(), batch, seed=0)
    pstep = GraphHSCNStep(_cfg(), batch, cuda, padded=False)
    _sync_weights(pstep, ostep)
    ostep.run()
    pstep.run_eager()
    got = pstep.download()
    torch.cuda.synchronize()
    want = torch.tensor(ostep.losses)
    assert_close(got[:2], want[:2], RTOL, "mincut / ortho loss")
    # the task loss goes through hard cluster ids: allow for near-tie flips of the fp32 softmax argmax
    assert rel_err(got[2:], want[2:]) < 1e-3
    for (n, a), (_, b) in zip(pstep.scn.named_parameters(), ostep.scn.named_parameters()):
        assert rel_err(a, b) < 1e-4, f"SCN param {n} after AdamW"
    for (n, a), (_, b) in zip(pstep.hscn.named_parameters(), ostep.hscn.named_parameters()):
        assert rel_err(a, b) < 2e-3, f"HSCN param {n} after AdamW"


def test_padded_captured_replay_equals_compact_eager(cuda):
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200._lib import lib
    from graph_hscn_b200.train import GraphHSCNStep
    batch = synthetic.peptides_batch(10, seed=78)
    a = GraphHSCNStep(_cfg(), batch, cuda, padded=False, seed=3)
    b = GraphHSCNStep(_cfg(), batch, cuda, padded=True, seed=3)
    _sync_weights(b, a)
    init = ({k: v.clone() for k, v in a.scn.state_dict().items()}, {k: v.clone() for k, v in a.hscn.state_dict().items()})
    b.capture(warmup=2)                       # warm-up steps move b's weights: reset both sides afterwards
    for st in (a, b):
        st.scn.load_state_dict(init[0])
        st.hscn.load_state_dict(init[1])
        for opt in (st.scn_opt, st.hscn_opt):
            for state in opt.state.values():
                for k, v in state.items():
                    if torch.is_tensor(v):
                        v.zero_()
    n0 = lib().query("ghscn_launch_count")
    for _ in range(3):
        a.run_eager()
        b.upload()
        b.run()
    la, lb = a.download().clone(), b.download().clone()
    torch.cuda.synchronize()
    assert lib().query("ghscn_launch_count") > n0
    assert_close(lb, la, 1e-4, "losses after 3 steps: captured padded vs eager compact")
    for (n, p), (_, q) in zip(a.hscn.named_parameters(), b.hscn.named_parameters()):
        assert rel_err(q, p) < 1e-3, f"HSCN param {n}"
    for (n, p), (_, q) in zip(a.scn.named_parameters(), b.scn.named_parameters()):
        assert rel_err(q, p) < 1e-4, f"SCN param {n}"


def test_parallel_branch_schedule_is_bit_identical(cuda):
    """The step runs the SCN stage + cluster assignment + the "virtual" HeteroConv branch on a side stream and the
    "local" half of the HSCN on the caller's stream (fork/join inside the CUDA graph): same kernels, same inputs ->
    parameters after several steps must equal the serial schedule bit for bit, eager and captured."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200 import train as gtrain
    from graph_hscn_b200.pyg import nn as pnn
    from graph_hscn_b200.train import GraphHSCNStep, StepConfig

    def run(parallel, captured):
        old = pnn.PARALLEL_BRANCHES, gtrain.TWO_STREAMS
        pnn.PARALLEL_BRANCHES = gtrain.TWO_STREAMS = parallel
        try:
            torch.manual_seed(0)
            step = GraphHSCNStep(StepConfig(hidden=64), synthetic.peptides_batch(16, seed=5), cuda, padded=True)
            n = 6
            if captured:
                step.capture(warmup=2)
                n -= 2
            for _ in range(n):
                step.run()
            torch.cuda.synchronize()
            return (torch.cat([step.hscn_grads.flat_param.detach(), step.scn_grads.flat_param.detach()]).clone(),
                    step.losses.clone())
        finally:
            pnn.PARALLEL_BRANCHES, gtrain.TWO_STREAMS = old

    for captured in (False, True):
        (p0, l0), (p1, l1) = run(False, captured), run(True, captured)
        assert torch.equal(p0, p1) and torch.equal(l0, l1), f"captured={captured}"
