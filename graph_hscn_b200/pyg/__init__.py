"""`torch_geometric` / `torch_scatter`-shaped namespace over the sm_100a kernels.

`install()` publishes it under the third-party module names the reference imports
(SURVEY.md 8b), so `graph_hscn/model/*.py` and `graph_hscn/train/*.py` run unchanged:

    import graph_hscn_b200.pyg as pyg; pyg.install()
    from graph_hscn.model.mpnn import MPNN       # resolves GCNConv, scatter_mean, ... to this package
"""
from __future__ import annotations

import sys
import types

from ..data import Batch, Data, DataLoader, HeteroBatch, HeteroData
from .. import ops as _ops
from . import _device, functional, nn
from ._device import auto_device, set_auto_device
from .functional import (SparseAdj, dense_mincut_pool, gcn_norm, global_add_pool, global_mean_pool,
                         mincut_pool_ragged, scatter, scatter_add, scatter_mean, scatter_sum, scn_logits_fused,
                         to_dense_adj)
from .nn import (GATConv, GCNConv, GINConv, GraphConv, HeteroConv, Linear, MessagePassing, Sequential)

__all__ = ["GATConv", "GCNConv", "GINConv", "GraphConv", "HeteroConv", "Linear", "MessagePassing", "Sequential",
           "dense_mincut_pool", "gcn_norm", "global_add_pool", "global_mean_pool", "mincut_pool_ragged", "scatter",
           "scatter_add", "scatter_mean", "scatter_sum", "to_dense_adj", "SparseAdj", "install", "namespace", "set_auto_device", "auto_device"]


def namespace() -> types.SimpleNamespace:
    """The operator set the mirror models (graph_hscn_b200.models) are parameterised with."""
    return types.SimpleNamespace(
        name="ghscn-b200", fused_relu=True, GCNConv=GCNConv, GATConv=GATConv, GINConv=GINConv, GraphConv=GraphConv,
        HeteroConv=HeteroConv, Linear=Linear, Sequential=Sequential, MessagePassing=MessagePassing,
        dense_mincut_pool=dense_mincut_pool, mincut_pool_ragged=mincut_pool_ragged, to_dense_adj=to_dense_adj,
        global_mean_pool=global_mean_pool, scatter_mean=scatter_mean, gcn_norm=gcn_norm,
        scn_logits_fused=scn_logits_fused, linear_act=functional.linear_act, relu_dropout=_ops.relu_dropout)


def build_modules(ns: types.SimpleNamespace, data_mod=None) -> dict:
    """Module objects mimicking the torch_geometric / torch_scatter import paths for operator set `ns`."""
    from .. import data as _data
    data_mod = data_mod or _data

    def mod(name: str, **attrs) -> types.ModuleType:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    conv_attrs = dict(GCNConv=ns.GCNConv, GATConv=ns.GATConv, GINConv=ns.GINConv, GraphConv=ns.GraphConv,
                      HeteroConv=ns.HeteroConv, MessagePassing=ns.MessagePassing)
    gcn_conv = mod("torch_geometric.nn.conv.gcn_conv", gcn_norm=ns.gcn_norm, GCNConv=ns.GCNConv)
    conv = mod("torch_geometric.nn.conv", gcn_conv=gcn_conv, **conv_attrs)
    conv.__path__ = []
    nn_mod = mod("torch_geometric.nn", conv=conv, Linear=ns.Linear, Sequential=ns.Sequential,
                 dense_mincut_pool=ns.dense_mincut_pool, global_mean_pool=ns.global_mean_pool, **conv_attrs)
    nn_mod.__path__ = []
    from . import utils as _utils          # transform/posenc.py:5-9 imports these three (+ remove_self_loops)
    utils = mod("torch_geometric.utils", to_dense_adj=ns.to_dense_adj,
                **{name: getattr(ns, name, getattr(_utils, name))
                   for name in ("get_laplacian", "to_undirected", "to_scipy_sparse_matrix", "remove_self_loops")})
    batch_mod = mod("torch_geometric.data.batch", Batch=data_mod.Batch)
    data = mod("torch_geometric.data", Data=data_mod.Data, HeteroData=data_mod.HeteroData, Batch=data_mod.Batch,
               batch=batch_mod)
    data.__path__ = []
    loader = mod("torch_geometric.loader", DataLoader=data_mod.DataLoader)
    root = mod("torch_geometric", nn=nn_mod, utils=utils, data=data, loader=loader, __version__="2.3.0+ghscn")
    root.__path__ = []
    scatter_mod = mod("torch_scatter", scatter_mean=ns.scatter_mean,
                      scatter=getattr(ns, "scatter", None), scatter_add=getattr(ns, "scatter_add", None))
    return {"torch_geometric": root, "torch_geometric.nn": nn_mod, "torch_geometric.nn.conv": conv,
            "torch_geometric.nn.conv.gcn_conv": gcn_conv, "torch_geometric.utils": utils,
            "torch_geometric.data": data, "torch_geometric.data.batch": batch_mod,
            "torch_geometric.loader": loader, "torch_scatter": scatter_mod}


def install(ns: types.SimpleNamespace = None, auto_device: bool = True) -> None:
    """Make `import torch_geometric...` / `import torch_scatter` resolve to this package.  `auto_device` (default on
    for the drop-in route) lets the operators accept the CPU tensors the reference hands them at
    train/train_clustering.py:37-43 (gcn_norm before `.to(device)`) and on its MPNN path (train/train.py:78-81):
    they are staged to the GPU, computed there and returned on the input's device (pyg/_device.py)."""
    if ns is None:
        set_auto_device(auto_device)
    ns = ns or namespace()
    if not hasattr(ns, "scatter"):
        ns.scatter, ns.scatter_add = scatter, scatter_add
    sys.modules.update(build_modules(ns))
