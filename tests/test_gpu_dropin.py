"""GPU tests of the drop-in route: `pyg.install()` + the reference's own call patterns, through the
`torch_geometric.*` / `torch_scatter` import names, checked against the CPU oracle running the same loop.

The reference sources are not on the GPU box, so the loops below are line-by-line transcriptions of
    train/train_clustering.py:34-69   (gcn_norm on CPU tensors BEFORE data.to(device); per-graph optimizer steps;
                                       clust.max(1)[1].cpu().numpy())
    train/train.py:73-95              (HSCN route: batch.to(device); MPNN route: model and batch stay on the CPU;
                                       criterion, backward, clip_grad_norm, optimizer.step)
with the models taken from graph_hscn_b200.models (pinned bit-for-bit to model/mpnn.py / model/hscn.py by
tests/test_golden_cpu.py).  fp32 tolerance 1e-5 relative on losses / predictions, integers bit-exact."""
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.util import RTOL, assert_close, copy_params, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def shim(cuda):
    from graph_hscn_b200 import pyg
    before = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith(("torch_geometric", "torch_scatter"))}
    pyg.install()
    assert pyg.auto_device()
    try:
        yield pyg
    finally:
        pyg.set_auto_device(False)
        for k in [k for k in sys.modules if k.startswith(("torch_geometric", "torch_scatter"))]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in before.items() if v is not None})


class FreshDataset:
    """InMemoryDataset semantics: every access hands out a fresh copy (train_clustering mutates `data`)."""

    def __init__(self, graphs):
        self.graphs = graphs

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        if i >= len(self.graphs):
            raise IndexError
        return self.graphs[i].clone()


def _train_clustering(dataset, model, gcn_norm, device, epochs=2, lr=1e-2):
    """train/train_clustering.py:28-69, logging stripped."""
    optimizer = torch.optim.AdamW(lr=lr, weight_decay=5e-4, params=model.parameters())
    losses = []
    for epoch in range(epochs):
        for data in dataset:
            data.edge_index, data.edge_weight = gcn_norm(
                data.edge_index,
                data.edge_weight,
                data.num_nodes,
                add_self_loops=True,
            )
            data = data.to(device)
            optimizer.zero_grad()
            _, mc_loss, o_loss, adj = model(
                data.x.float(), data.edge_index, data.edge_weight
            )
            loss = mc_loss + o_loss
            loss.backward()
            optimizer.step()
            losses.append(loss.item())
    cluster_all_lst, softs = [], []
    for data in dataset:
        data.edge_index, data.edge_weight = gcn_norm(
            data.edge_index,
            data.edge_weight,
            data.num_nodes,
            add_self_loops=True,
        )
        data = data.to(device)
        clust, _, _, adj = model(
            data.x.float(), data.edge_index, data.edge_weight
        )
        clusters = clust.max(1)[1].cpu().numpy()
        cluster_all_lst.append(clusters)
        softs.append(clust.detach().cpu())
    return cluster_all_lst, softs, losses


def test_train_clustering_call_pattern_on_the_cuda_shim(shim):
    """gcn_norm receives CPU tensors (staged to the GPU, results back on the CPU), then everything moves to CUDA."""
    from torch_geometric.nn.conv.gcn_conv import gcn_norm
    import torch_geometric.nn as tgnn
    from graph_hscn_b200 import models, synthetic
    from graph_hscn_b200._lib import lib
    from oracle.namespace import namespace as oracle_ns
    assert tgnn.GraphConv is shim.GraphConv and tgnn.dense_mincut_pool is shim.dense_mincut_pool
    graphs = synthetic.peptides_graphs(6, seed=41)
    K = 10
    torch.manual_seed(0)
    ref = models.SCN([16], "elu", 9, K, ops=oracle_ns())
    got = models.SCN([16], "elu", 9, K).to("cuda")
    copy_params(got, ref)
    n0 = lib().launches
    c_got, s_got, l_got = _train_clustering(FreshDataset(graphs), got, gcn_norm, torch.device("cuda"))
    assert lib().launches > n0, "the CUDA library was not used"
    c_ref, s_ref, l_ref = _train_clustering(FreshDataset(graphs), ref, oracle_ns().gcn_norm, torch.device("cpu"))
    # 12 optimizer steps apart the losses still agree to fp32 noise
    assert rel_err(torch.tensor(l_got), torch.tensor(l_ref)) < 1e-4
    for (n, a), (_, b) in zip(got.named_parameters(), ref.named_parameters()):
        assert rel_err(a, b) < 1e-3, f"SCN param {n} after the clustering loop"
    agree = total = 0
    for cg, cr, sr in zip(c_got, c_ref, s_ref):
        assert isinstance(cg, np.ndarray) and cg.dtype == cr.dtype
        top2 = sr.topk(2, dim=1)[0]
        safe = ((top2[:, 0] - top2[:, 1]) > 1e-3).numpy()
        assert np.array_equal(cg[safe], cr[safe])
        agree += int((cg == cr).sum())
        total += cr.size
    assert agree / total > 0.99


def test_gcn_norm_on_cpu_tensors_is_bit_exact_and_returns_cpu(shim):
    from torch_geometric.nn.conv.gcn_conv import gcn_norm
    from graph_hscn_b200 import synthetic
    from oracle import ops as oops
    d = synthetic.peptides_graphs(1, seed=3)[0]
    ei, ew = gcn_norm(d.edge_index, d.edge_weight, d.num_nodes, add_self_loops=True)
    assert ei.device.type == "cpu" and ew.device.type == "cpu"
    ei_r, ew_r = oops.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
    assert torch.equal(ei, ei_r) and torch.equal(ew, ew_r)


def _hetero_list(graphs, K, seed):
    from oracle import hetero as ohet
    from torch_geometric.data import HeteroData
    rng = np.random.default_rng(seed)
    out = []
    for d in graphs:
        clusters = rng.integers(0, K, size=d.num_nodes)
        _, vx, vv, lv = ohet.virtual_nodes(d.x, clusters, K)
        h = HeteroData()
        h["local"].x, h["local"].y, h["virtual"].x = d.x.float(), d.y, vx
        h["local", "to", "local"].edge_index = d.edge_index
        h["virtual", "to", "virtual"].edge_index = vv
        h["local", "to", "virtual"].edge_index = lv
        out.append(h)
    return out


def _train_epoch(loader, model, optimizer, loss_fn, is_hscn, device, batch_accumulation=1, clip_grad_norm=True):
    """train/train.py:53-95 (metrics / logging stripped)."""
    from graph_hscn_b200.models import criterion
    model.train()
    optimizer.zero_grad()
    losses, preds = [], []
    for _iter, batch in enumerate(loader):
        if is_hscn:
            batch = batch.to(device)
            pred = model(batch.x_dict, batch.edge_index_dict, batch)
            true = batch["local"].y
        else:
            batch.x = batch.x.float()
            pred = model(batch)
            true = batch.y
        loss, pred_score = criterion(loss_fn, pred, true)
        losses.append(loss.item())
        preds.append(pred.detach().cpu())
        loss.backward()
        if ((_iter + 1) % batch_accumulation == 0) or (_iter + 1 == len(loader)):
            if clip_grad_norm:
                nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            optimizer.step()
            optimizer.zero_grad()
    return losses, preds


def test_train_epoch_hscn_route_on_the_cuda_shim(shim):
    from torch_geometric.loader import DataLoader
    from graph_hscn_b200 import models, synthetic
    from oracle.namespace import namespace as oracle_ns
    graphs = synthetic.peptides_graphs(12, seed=43, task="struct")
    hl = _hetero_list(graphs, 10, seed=5)
    torch.manual_seed(1)
    ref = models.HSCN("GAT", "GCN", "GCN", F.relu, 9, 40, 11, 3, ops=oracle_ns())
    first = next(iter(DataLoader(hl, batch_size=4)))
    ref(first.x_dict, first.edge_index_dict, first)                      # materialise lazy parameters
    got = models.HSCN("GAT", "GCN", "GCN", F.relu, 9, 40, 11, 3).to("cuda")
    copy_params(got, ref)
    kw = dict(lr=1e-3, weight_decay=5e-4)
    l_ref, p_ref = _train_epoch(DataLoader(hl, batch_size=4), ref, torch.optim.AdamW(ref.parameters(), **kw), "l1",
                                True, torch.device("cpu"), batch_accumulation=2)
    l_got, p_got = _train_epoch(DataLoader(hl, batch_size=4), got, torch.optim.AdamW(got.parameters(), **kw), "l1",
                                True, torch.device("cuda"), batch_accumulation=2)
    assert_close(torch.tensor(l_got), torch.tensor(l_ref), 1e-4, "losses over 3 variable-shape batches")
    assert_close(p_got[0], p_ref[0], RTOL, "first prediction")
    for (n, a), (_, b) in zip(got.named_parameters(), ref.named_parameters()):
        assert rel_err(a, b) < 1e-3, f"HSCN param {n} after the epoch"


def test_train_epoch_mpnn_route_stays_on_cpu_tensors(shim):
    """train.py:78-81 / main.py:117: the MPNN and its batches are never moved -- CPU tensors in, GPU compute, CPU out."""
    from torch_geometric.loader import DataLoader
    from torch_geometric.nn import GCNConv
    from graph_hscn_b200 import models, synthetic
    from graph_hscn_b200._lib import lib
    from oracle.namespace import namespace as oracle_ns
    graphs = synthetic.peptides_graphs(10, seed=44)
    torch.manual_seed(2)
    ref = models.MPNN(oracle_ns().GCNConv, F.relu, 9, 32, 10, 4, ops=oracle_ns())
    got = models.MPNN(GCNConv, F.relu, 9, 32, 10, 4)                     # stays on the CPU, like build_mpnn's result
    copy_params(got, ref)
    assert all(not p.is_cuda for p in got.parameters())
    kw = dict(lr=1e-3, weight_decay=5e-4)
    n0 = lib().launches
    l_got, p_got = _train_epoch(DataLoader(graphs, batch_size=5), got, torch.optim.AdamW(got.parameters(), **kw),
                                "cross_entropy", False, torch.device("cuda"))
    assert lib().launches > n0, "the CUDA library was not used"
    l_ref, p_ref = _train_epoch(DataLoader(graphs, batch_size=5), ref, torch.optim.AdamW(ref.parameters(), **kw),
                                "cross_entropy", False, torch.device("cpu"))
    assert p_got[0].device.type == "cpu"
    assert_close(p_got[0], p_ref[0], RTOL, "first prediction")
    assert_close(torch.tensor(l_got), torch.tensor(l_ref), 1e-4, "losses")
    for (n, a), (_, b) in zip(got.named_parameters(), ref.named_parameters()):
        assert not a.is_cuda
        assert rel_err(a, b) < 1e-3, f"MPNN param {n} after the epoch"


def test_cpu_tensor_still_raises_without_auto_device(cuda):
    from graph_hscn_b200 import pyg, synthetic
    assert not pyg.auto_device()
    d = synthetic.peptides_graphs(1, seed=3)[0]
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pyg.gcn_norm(d.edge_index, None, d.num_nodes)
