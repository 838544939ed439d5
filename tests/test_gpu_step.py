"""GPU tests of the whole Graph-HSCN step: CUDA step vs the CPU oracle step, padded vs compact virtual
layout, CUDA-graph replay vs eager execution."""
import pytest
import torch

from tests.util import RTOL, assert_close, rel_err

pytestmark = pytest.mark.gpu


def _cfg():
    from graph_hscn_b200.train import StepConfig
    return StepConfig(hidden=48, num_layers=3, num_clusters=10, lr=1e-2)


def _read_losses(step):
    """`download()` is an asynchronous D2H copy into pinned memory: wait for it before looking."""
    host = step.download()
    torch.cuda.synchronize()
    return host.clone()


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
    got = _read_losses(pstep)
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
    la, lb = _read_losses(a), _read_losses(b)
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


# ------------------------------------------------------------------------------------------------ round 2
def _flat_named(grads):
    return dict(zip(grads.names, (p.grad for p in grads.params)))


def test_bench_configuration_step_matches_oracle(cuda):
    """The benchmark's own configuration -- StepConfig() (h=300, K=10, 3 layers), the Peptides-func batch of seed 1236
    (128 graphs), bucketed padding, ONE captured two-stream CUDA graph -- against the CPU oracle step on the same
    batch and weights: losses, HSCN logits, live gradients of both models, cluster-id agreement."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig
    from oracle.step import OracleStep
    batch = synthetic.peptides_batch(128, seed=1236, task="func")
    cfg = StepConfig()
    ostep = OracleStep(cfg, batch, seed=0)
    pstep = GraphHSCNStep(cfg, batch, cuda, padded=True, policy=BucketPolicy.for_batches([batch]))
    assert pstep.runner.shape.dummies >= 1 and pstep.runner.shape.n_cap > batch.x.size(0)
    _sync_weights(pstep, ostep)
    pred0 = pstep.predict()                       # HSCN logits before any update (independent of the clusters)
    pstep.capture(warmup=0)
    _sync_weights(pstep, ostep)                   # (the dry run is rolled back; load again to be explicit)
    pstep.run()
    got = _read_losses(pstep)
    ostep.run()
    want = torch.tensor(ostep.losses)
    assert_close(got[:2], want[:2], RTOL, "mincut / ortho loss at the bench configuration")
    assert_close(got[2:], want[2:], RTOL, "task loss at the bench configuration")
    assert_close(pred0, ostep.pred, RTOL, "HSCN logits at the bench configuration")
    for n, g in _flat_named(pstep.scn_grads).items():
        assert_close(g, ostep.scn_grads[n], 1e-4, f"SCN gradient {n}")
    live = _flat_named(pstep.hscn_grads)
    assert set(live) == set(ostep.grads)
    for n, g in live.items():
        assert_close(g, ostep.grads[n], 1e-4, f"HSCN gradient {n}")
    # cluster ids: bit-exact outside fp32 near-ties of the softmax; report the near-tie rate
    from graph_hscn_b200 import hetero
    from graph_hscn_b200.structure import structure_hints
    b = batch
    ei, ew = ostep.ns.gcn_norm(b.edge_index, None, b.x.size(0), add_self_loops=True)
    with torch.no_grad():
        s_ref = torch.softmax(ostep.scn.logits(b.x.float(), ei, ew), dim=-1)
    N = b.x.size(0)
    with torch.no_grad(), structure_hints(**pstep.hints):
        d = pstep.dev
        x_f = pstep._cast(d["x"])
        ei_d, ew_d = pstep.ns.gcn_norm(d["edge_index"], None, x_f.size(0), add_self_loops=True)
        ids = hetero.assign_clusters(torch.softmax(pstep.scn.logits(x_f, ei_d, ew_d), dim=-1))[:N].cpu().long()
    want_ids = s_ref.max(1)[1]
    top2 = s_ref.topk(2, dim=1)[0]
    near = (top2[:, 0] - top2[:, 1]) <= 1e-5
    assert torch.equal(ids[~near], want_ids[~near])
    rate, agree = float(near.float().mean()), float((ids == want_ids).float().mean())
    print(f"bench configuration: near-tie rate {rate:.2e}, cluster-id agreement {agree:.6f} over {N} nodes")
    assert agree > 0.999


def test_bucketed_captured_steps_equal_exact_shape_eager_steps(cuda):
    """Four different-sized batches through (a) exact shapes, eager, compact virtual layout and (b) bucketed padding
    with dummy graphs, one captured graph per bucket: same losses every step and same weights at the end."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep
    batches = [synthetic.peptides_batch(10, seed=s) for s in (31, 32, 33, 31)]
    pol = BucketPolicy.for_batches(batches, node_step=128, edge_step=256)
    a = GraphHSCNStep(_cfg(), batches[0], cuda, padded=False, seed=3)
    b = GraphHSCNStep(_cfg(), batches[0], cuda, padded=True, seed=3, policy=pol, auto_capture=True)
    _sync_weights(b, a)
    for i, bt in enumerate(batches):
        a.load(bt)
        a.run()
        b.load(bt)
        b.run()
        la, lb = _read_losses(a), _read_losses(b)
        assert_close(lb[:2], la[:2], 2e-5, f"SCN losses, step {i}")
        assert rel_err(lb[2:], la[2:]) < 1e-4, f"task loss, step {i}"
    assert 2 <= b.num_buckets <= 3 and b.num_graphs_captured == b.num_buckets     # batch 4 replays bucket 1's graph
    for (n, p), (_, q) in zip(a.hscn.named_parameters(), b.hscn.named_parameters()):
        assert rel_err(q, p) < 1e-3, f"HSCN param {n}"
    for (n, p), (_, q) in zip(a.scn.named_parameters(), b.scn.named_parameters()):
        assert rel_err(q, p) < 1e-4, f"SCN param {n}"


def test_accumulation_and_clip_match_oracle_loop(cuda):
    """train/train.py:87-95: gradients accumulate over `batch_accumulation` batches, are clipped to norm 1 and applied;
    captured bucketed steps vs the CPU oracle's loop over the same three batches."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig
    from oracle.step import OracleStep
    cfg = StepConfig(hidden=48, lr=1e-2, batch_accumulation=2, clip_grad_norm=True, loss_fn="l1", num_classes=11)
    batches = [synthetic.peptides_batch(8, seed=s, task="struct") for s in (51, 52, 53, 54)]
    ostep = OracleStep(cfg, batches[0], seed=0)
    pstep = GraphHSCNStep(cfg, batches[0], cuda, padded=True, policy=BucketPolicy.for_batches(batches, 128, 256),
                          auto_capture=True)
    _sync_weights(pstep, ostep)
    for i, bt in enumerate(batches):
        ostep.set_batch(bt)
        ostep.run()
        pstep.load(bt)
        pstep.run()
        got = _read_losses(pstep)
        assert_close(got[:2], torch.tensor(ostep.losses[:2]), 1e-4, f"SCN losses, step {i}")
        assert rel_err(got[2:], torch.tensor(ostep.losses[2:])) < 1e-3, f"task loss, step {i}"
    for (n, a), (_, b) in zip(pstep.hscn.named_parameters(), ostep.hscn.named_parameters()):
        assert rel_err(a, b) < 2e-3, f"HSCN param {n} after two clipped, accumulated updates"


def test_grad_clip_kernel_matches_torch(cuda):
    from graph_hscn_b200.train import FlatAdamW
    g = torch.Generator().manual_seed(3)
    for scale in (1e-3, 10.0):
        grad = (torch.randn(276_910, generator=g) * scale).to(cuda)
        ref = torch.nn.Parameter(torch.zeros_like(grad))
        ref.grad = grad.clone()
        total = torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt = FlatAdamW(torch.zeros_like(grad), grad.clone())
        out = opt.clip_grad_norm(1.0)
        assert_close(out[0:1], total.view(1), 1e-6, "total norm")
        assert_close(opt.g * out[1], ref.grad, 1e-6, "clipped gradient")


def test_prefetch_pipeline_delivers_the_right_batch(cuda):
    """Double-buffered H2D on the copy stream: the bucket's static buffer holds exactly the staged bytes."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep
    batches = [synthetic.peptides_batch(6, seed=s) for s in (61, 62, 63)]
    step = GraphHSCNStep(_cfg(), batches[0], cuda, policy=BucketPolicy.for_batches(batches, 64, 128))
    staged = [step.stage(b) for b in batches]
    slot = step.prefetch(staged[0])
    for i, st in enumerate(staged):
        step.select_prefetched(st, slot)
        if i + 1 < len(staged):
            slot = step.prefetch(staged[i + 1])
        torch.cuda.synchronize()
        assert torch.equal(step.runner.dev_buf.cpu(), st.buf)
        assert torch.equal(step.dev["x"][:st.num_nodes].cpu(), batches[i].x)


def test_device_collate_equals_host_collate(cuda):
    """loader/loader.py:48-60 on the device: graphs packed with LOCAL edge indices + per-graph counts, `batch` and the
    offset edge_index derived by ghscn_collate_batch -- bit-identical to Batch.from_data_list, and the step computes
    the same losses from either staging."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StagingCollate
    graphs = synthetic.peptides_graphs(10, seed=71)
    host = Batch.from_data_list(graphs)
    pol = BucketPolicy.for_batches([host], node_step=128, edge_step=256)
    a = GraphHSCNStep(_cfg(), host, cuda, padded=True, seed=3, policy=pol)
    b = GraphHSCNStep(_cfg(), host, cuda, padded=True, seed=3, policy=pol)
    _sync_weights(b, a)
    raw = StagingCollate(pol, pin=True)(graphs)
    assert raw.raw and raw.nbytes < a.staged.nbytes                  # `batch` is not uploaded
    b.load(raw)
    b.predict()                                                      # runs the device collate
    torch.cuda.synchronize()
    N, E = host.x.size(0), host.edge_index.size(1)
    assert torch.equal(b.dev["batch"].cpu(), a.staged.views["batch"])
    assert torch.equal(b.dev["edge_index"].cpu(), a.staged.views["edge_index"])
    assert torch.equal(b.dev["edge_index"][:, :E].cpu(), host.edge_index) and torch.equal(b.dev["batch"][:N].cpu(), host.batch)
    for st in (a, b):
        st.capture(warmup=0)
    for _ in range(2):
        a.run()
        b.run()
    la, lb = _read_losses(a), _read_losses(b)
    assert torch.equal(la, lb), "device-collated and host-collated batches must give bit-identical steps"


def test_staging_collate_in_a_dataloader(cuda):
    """`DataLoader(dataset, batch_size, collate_fn=StagingCollate(policy))` feeds the bucketed step directly."""
    from torch.utils.data import DataLoader as TorchLoader
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StagingCollate
    graphs = synthetic.peptides_graphs(24, seed=72)
    pol = BucketPolicy(444, 1024, node_step=128, edge_step=256)
    step = GraphHSCNStep(_cfg(), Batch.from_data_list(graphs[:8]), cuda, policy=pol, auto_capture=True)
    seen = 0
    for staged in TorchLoader(graphs, batch_size=8, collate_fn=StagingCollate(pol, pin=True)):
        step.load(staged)
        step.run()
        losses = _read_losses(step)
        assert torch.isfinite(losses).all()
        seen += 1
    assert seen == 3 and step.num_graphs_captured >= 1


def test_overlapped_gradient_exchange_equals_plain_backward(cuda):
    """Data parallel: the all-reduce of everything behind the first HSCN layer is issued while that layer's
    backward still runs (two autograd passes split at the layer's output).  On a one-rank NCCL group the flat gradient
    buffer must equal the one of the plain backward bit for bit."""
    import torch.distributed as dist
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.structure import structure_cache, structure_hints
    from graph_hscn_b200.train import GraphHSCNStep
    if dist.is_initialized():
        pytest.skip("a process group is already active")
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29547", rank=0, world_size=1)
    try:
        step = GraphHSCNStep(_cfg(), synthetic.peptides_batch(12, seed=5), cuda, padded=True)
        g = step.hscn_grads
        split = g.leading("convs.0.")
        assert 0 < split < len(g.params) and g.leading("lin_") == 0
        flats = []
        for overlapped in (False, True):
            with structure_hints(**step.hints):
                structure_cache().clear()
                step._register_blocks()
                x_f = step._cast(step.dev["x"])
                ei, ew, *_ = step._forward_scn(x_f)
                hb = step._assign(x_f, ei, ew)
                loss = step._hscn_loss(hb)
                boundary = step.hscn.first_layer_output
                assert boundary is not None and boundary.requires_grad
                g.flat.fill_(float("nan"))
                if overlapped:
                    g.backward_reduce_overlapped(loss, boundary, split)
                else:
                    g.backward_into(loss)
            torch.cuda.synchronize()
            flats.append(g.flat.clone())
        assert torch.isfinite(flats[0]).all() and torch.equal(flats[0], flats[1])
    finally:
        dist.destroy_process_group()
