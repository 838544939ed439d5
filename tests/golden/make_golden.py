"""Generates tests/golden/*.pt by running the UNMODIFIED reference sources on top of the CPU oracle.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What is pinned: the reference's own model / training / data code
    graph_hscn/model/mpnn.py, graph_hscn/model/hscn.py,
    graph_hscn/train/train_clustering.py, graph_hscn/loader/hetero_data.py
imported unchanged from /root/reference, with `torch_geometric` / `torch_scatter` resolving to the
oracle restatement (oracle/ops.py, oracle/nn.py) and to this repo's Data/Batch containers.  What is NOT
pinned (PyG is not installable offline): the oracle operators against a real PyG build -- see
oracle/__init__.py, "parity unpinned".
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)


def install_reference_imports() -> None:
    import pydantic.v1 as pyd_v1
    from graph_hscn_b200 import pyg
    from oracle.namespace import namespace
    sys.modules["pydantic"] = pyd_v1                       # config/config.py uses the pydantic-v1 API
    sys.modules.update(pyg.build_modules(namespace()))     # torch_geometric.*, torch_scatter -> oracle
    ogb = types.ModuleType("ogb")
    gpp = types.ModuleType("ogb.graphproppred")
    gpp.PygGraphPropPredDataset = object
    ogb.graphproppred = gpp
    sys.modules["ogb"], sys.modules["ogb.graphproppred"] = ogb, gpp
    wandb = types.ModuleType("wandb")
    wandb.log = lambda *a, **k: None
    sys.modules["wandb"] = wandb
    loader_stub = types.ModuleType("graph_hscn.loader.loader")   # avoids rdkit/ogb dataset imports
    loader_stub.get_loader = lambda *a, **k: None
    sys.path.insert(0, REFERENCE)
    import graph_hscn.loader  # noqa: F401
    sys.modules["graph_hscn.loader.loader"] = loader_stub


class FreshDataset:
    """InMemoryDataset semantics: every access hands out a fresh copy (train_clustering mutates `data`)."""

    def __init__(self, graphs):
        self.graphs = graphs

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        if isinstance(i, (list, torch.Tensor, np.ndarray)):
            return FreshDataset([self.graphs[int(j)] for j in i])
        return self.graphs[int(i)].clone()

    def __iter__(self):
        return (g.clone() for g in self.graphs)


class QuietLogger:
    def info(self, *_a, **_k):
        pass


def main() -> None:
    install_reference_imports()
    import torch.nn.functional as F
    from graph_hscn.loader.hetero_data import generate_hetero_data
    from graph_hscn.loss import criterion
    from graph_hscn.model.hscn import HSCN, SCN
    from graph_hscn.model.mpnn import MPNN
    from graph_hscn.train.train_clustering import train_clustering
    from torch_geometric.nn import GCNConv

    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    SNS = types.SimpleNamespace
    out = {}

    # ---- config #1 shape: MPNN(GCN) ------------------------------------------------------------
    graphs = synthetic.peptides_graphs(6, seed=101, task="func")
    batch = Batch.from_data_list(graphs)
    batch.x = batch.x.float()                                  # train.py:79
    torch.manual_seed(11)
    mpnn = MPNN(GCNConv, F.relu, 9, 32, 10, 4, dropout=0.0)
    mpnn.eval()
    pred = mpnn(batch)
    loss, score = criterion("cross_entropy", pred, batch.y)
    loss.backward()
    out["mpnn"] = dict(
        graphs=[dict(x=g.x, edge_index=g.edge_index, y=g.y) for g in graphs], state=mpnn.state_dict(),
        pred=pred.detach(), loss=loss.detach(), score=score.detach(),
        grads={n: p.grad.clone() for n, p in mpnn.named_parameters()})

    # ---- SCN + train_clustering (reference loop, per graph) ---------------------------------------
    K = 5
    cgraphs = synthetic.peptides_graphs(5, seed=202, task="struct")
    torch.manual_seed(22)
    scn = SCN([16], "elu", 9, K)
    init_state = {k: v.clone() for k, v in scn.state_dict().items()}
    clusters = train_clustering(QuietLogger(), FreshDataset(cgraphs), scn, SNS(cluster_epochs=2),
                                SNS(optim_type="adamW", lr=0.01, weight_decay=5e-4), SNS(use_wandb=False))
    from torch_geometric.nn.conv.gcn_conv import gcn_norm
    g0 = cgraphs[0].clone()
    ei, ew = gcn_norm(g0.edge_index, None, g0.num_nodes, add_self_loops=True)
    S, mc, ol, adj = scn(g0.x.float(), ei, ew)
    out["scn"] = dict(
        graphs=[dict(x=g.x, edge_index=g.edge_index, y=g.y) for g in cgraphs], K=K, init_state=init_state,
        final_state={k: v.clone() for k, v in scn.state_dict().items()},
        clusters=[torch.from_numpy(np.asarray(c)).long() for c in clusters],
        g0=dict(S=S.detach(), mc=mc.detach(), ortho=ol.detach(), adj=adj.detach(), edge_index=ei, edge_weight=ew))

    # ---- generate_hetero_data (reference, per graph) + HSCN ------------------------------------------
    split = {"train": torch.arange(0, 3), "val": torch.arange(3, 4), "test": torch.arange(4, 5)}
    hlist = generate_hetero_data(clusters, FreshDataset(cgraphs), split, SNS(task_level="graph"),
                                 SNS(num_clusters=K), QuietLogger())
    out["hetero"] = [dict(local_x=h["local"].x, y=h["local"].y, virtual_x=h["virtual"].x,
                          ll=h["local", "to", "local"].edge_index, vv=h["virtual", "to", "virtual"].edge_index,
                          lv=h["local", "to", "virtual"].edge_index) for h in hlist]
    hb = Batch.from_data_list(hlist)
    torch.manual_seed(33)
    hscn = HSCN("GAT", "GCN", "GCN", F.relu, 9, 24, 11, 2)
    pred = hscn(hb.x_dict, hb.edge_index_dict, hb)
    loss, _ = criterion("l1", pred, hb["local"].y)
    loss.backward()
    out["hscn"] = dict(state=hscn.state_dict(), pred=pred.detach(), loss=loss.detach(),
                       grads={n: (p.grad.clone() if p.grad is not None else None) for n, p in hscn.named_parameters()})

    for name, blob in out.items():
        torch.save(blob, os.path.join(HERE, f"{name}.pt"))
        print(name, os.path.getsize(os.path.join(HERE, f"{name}.pt")), "bytes")


if __name__ == "__main__":
    main()
