"""Operator namespace of the CPU oracle (same attribute set as graph_hscn_b200.pyg.namespace()).

TEST INFRASTRUCTURE (see oracle/__init__.py): lets tests and bench.py's cpu_baseline run the mirror
models, or the reference's unchanged sources, on the pure-torch restatement of the PyG operators.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

from torch import Tensor

from . import nn as onn
from . import ops as oops
from . import posenc as oposenc


def mincut_pool_ragged(x: Tensor, edge_index: Tensor, s: Tensor, batch: Optional[Tensor] = None,
                       edge_attr: Optional[Tensor] = None, temp: float = 1.0, want_out: bool = True,
                       want_adj: bool = True):
    """PyG's batched recipe: to_dense_batch + to_dense_adj(batch) + dense_mincut_pool(mask)."""
    if batch is None:
        adj = oops.to_dense_adj(edge_index, None, edge_attr, max_num_nodes=x.size(0))
        return oops.dense_mincut_pool(x, adj, s, None, temp)
    xd, mask = oops.to_dense_batch(x, batch)
    sd, _ = oops.to_dense_batch(s, batch)
    adj = oops.to_dense_adj(edge_index, batch, edge_attr, max_num_nodes=xd.size(1))
    return oops.dense_mincut_pool(xd, adj, sd, mask, temp)


def namespace() -> SimpleNamespace:
    return SimpleNamespace(
        name="oracle-cpu", GCNConv=onn.GCNConv, GATConv=onn.GATConv, GINConv=onn.GINConv, GraphConv=onn.GraphConv,
        HeteroConv=onn.HeteroConv, Linear=onn.Linear, Sequential=onn.Sequential, MessagePassing=onn.MessagePassing,
        dense_mincut_pool=oops.dense_mincut_pool, mincut_pool_ragged=mincut_pool_ragged,
        to_dense_adj=oops.to_dense_adj, global_mean_pool=oops.global_mean_pool, scatter_mean=oops.scatter_mean,
        gcn_norm=oops.gcn_norm, scatter=oops.scatter, scatter_add=oops.scatter_add,
        get_laplacian=oposenc.get_laplacian, to_undirected=oposenc.to_undirected,
        to_scipy_sparse_matrix=oposenc.to_scipy_sparse_matrix, remove_self_loops=oposenc.remove_self_loops)
