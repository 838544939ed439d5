"""Data-parallel host logic on CPU: world_size 2, gloo.  Shards graphs by rank, one flat-buffer all-reduce
per model, and checks averaged gradients == single-process full-batch gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import rel_err


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(seed=0):
    from graph_hscn_b200 import models
    from oracle.namespace import namespace
    torch.manual_seed(seed)
    return models.MPNN("gcn", torch.relu, 9, 24, 10, 3, ops=namespace())


def _loss_on(model, graphs):
    from graph_hscn_b200 import models
    from graph_hscn_b200.data import Batch
    b = Batch.from_data_list(graphs)
    b.x = b.x.float()
    loss, _ = models.criterion("cross_entropy", model(b), b.y)
    return loss


def _worker(rank, world, port, q, uneven=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from graph_hscn_b200 import synthetic
        from graph_hscn_b200.train import FlatGradients, shard_range
        graphs = synthetic.peptides_graphs(8, seed=5)
        model = _build()
        flat = FlatGradients(model)
        if uneven:      # node-balanced cut points: the ranks hold different graph counts (5 + 3 here)
            lo, hi = (0, 5) if rank == 0 else (5, 8)
            flat.zero()
            _loss_on(model, graphs[lo:hi]).backward()
            flat.all_reduce_mean(local_graphs=hi - lo, global_graphs=len(graphs))
            q.put((rank, flat.flat.clone(), (lo, hi)))
            return
        lo, hi = shard_range(len(graphs), rank, world)
        flat.zero()
        _loss_on(model, graphs[lo:hi]).backward()
        assert all(p.grad.data_ptr() == flat.flat[o:o + 1].data_ptr()
                   for p, o in zip(flat.params, _offsets(flat.params)))   # grads live inside the flat buffer
        flat.all_reduce_mean()
        q.put((rank, flat.flat.clone(), (lo, hi)))
    finally:
        dist.destroy_process_group()


def _offsets(params):
    off, out = 0, []
    for p in params:
        out.append(off)
        off += p.numel()
    return out


def test_two_rank_allreduce_equals_full_batch():
    from graph_hscn_b200 import synthetic
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    assert got[0][2] == (0, 4) and got[1][2] == (4, 8)
    assert torch.equal(got[0][1], got[1][1])                # both ranks hold the same averaged gradient
    graphs = synthetic.peptides_graphs(8, seed=5)
    model = _build()
    _loss_on(model, graphs).backward()                      # BCE mean over 8 graphs == mean of the two shard means
    full = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert rel_err(got[0][1], full) < 1e-5


def test_two_rank_allreduce_with_uneven_shards_weights_by_graph_count():
    """Node-balanced shards hold different graph counts: each rank's batch-mean gradient is weighted by its share of
    the global graph count before the SUM all-reduce (a plain mean over ranks would be wrong)."""
    from graph_hscn_b200 import synthetic
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    assert got[0][2] == (0, 5) and got[1][2] == (5, 8)
    assert torch.equal(got[0][1], got[1][1])
    graphs = synthetic.peptides_graphs(8, seed=5)
    model = _build()
    _loss_on(model, graphs).backward()
    full = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert rel_err(got[0][1], full) < 1e-5


def test_flat_gradients_skip_dead_parameters():
    from graph_hscn_b200.train import FlatGradients, live_parameter_names
    lin_a, lin_b = torch.nn.Linear(3, 2), torch.nn.Linear(3, 2)
    m = torch.nn.ModuleDict({"a": lin_a, "b": lin_b})
    x = torch.randn(4, 3)
    live = live_parameter_names(m, lin_a(x).sum())
    assert live == ["a.weight", "a.bias"]
    flat = FlatGradients(m, live)
    assert flat.flat.numel() == 8 and lin_b.weight.grad is None
    flat.zero()
    lin_a(x).sum().backward()
    assert torch.allclose(flat.flat[:6].view(2, 3), x.sum(0).expand(2, 3))
