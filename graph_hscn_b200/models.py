"""Host-side mirror of the reference's model callers, parameterised by an operator namespace.

The reference's own model files (graph_hscn/model/mpnn.py, graph_hscn/model/hscn.py) run unchanged on
top of `graph_hscn_b200.pyg.install()`; they are not available on the GPU box, so the benchmark and the
GPU parity tests use these mirrors instead.  Same constructor arguments, same submodule / parameter
names (state_dicts interchange with the reference classes), same forward semantics including the
reference's quirks.  tests/golden pins mirror == unchanged reference source.

`ops` is an operator namespace: `graph_hscn_b200.pyg.namespace()` (CUDA kernels, default) or
`oracle.namespace()` (CPU oracle; only bench.py's cpu_baseline and tests pass that one in).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

ACTIVATIONS: Dict[str, Callable] = {   # config/config.py:13-18
    "elu": F.elu, "relu": F.relu, "tanh": torch.tanh, "identity": nn.Identity(),
}


def _default_ops() -> SimpleNamespace:
    from . import pyg
    return pyg.namespace()


class MPNN(nn.Module):
    """Stack of `conv` layers + mean readout (model/mpnn.py:13-60)."""

    def __init__(self, conv, activation: Callable, num_features: int, hidden_channels: int, num_classes: int,
                 num_layers: int, dropout: float = 0.0, use_batch_norm: bool = False,
                 use_layer_norm: bool = False, ops: Optional[SimpleNamespace] = None):
        super().__init__()
        self.ops = ops or _default_ops()
        if isinstance(conv, str):
            conv = {"gcn": self.ops.GCNConv, "gat": self.ops.GATConv, "gin": self.ops.GINConv}[conv.lower()]
        widths = [num_features] + [hidden_channels] * (num_layers - 1) + [num_classes]
        self.num_layers = num_layers
        self.conv_layers = nn.ModuleList(conv(widths[i], widths[i + 1]) for i in range(num_layers))
        self.use_batch_norm, self.use_layer_norm = use_batch_norm, use_layer_norm
        if use_layer_norm:  # the reference creates BOTH lists under this flag (Appendix B-8)
            self.bns = nn.ModuleList(nn.BatchNorm1d(hidden_channels) for _ in range(num_layers - 1))
            self.lns = nn.ModuleList(nn.LayerNorm(hidden_channels) for _ in range(num_layers - 1))
        self.activation, self.dropout = activation, dropout
        # operator sets that can run the ReLU of `F.relu(conv(x))` inside the aggregation kernel (same values)
        self._fused = bool(getattr(self.ops, "fused_relu", False)) and all(
            hasattr(c, "fuse_relu") for c in self.conv_layers[:-1])
        if self._fused:
            for c in self.conv_layers[:-1]:
                c.fuse_relu = True

    def forward(self, batch) -> Tensor:
        h, edge_index, graph_of_node = batch.x, batch.edge_index, batch.batch
        for i, conv in enumerate(self.conv_layers[:-1]):
            h = conv(h, edge_index) if self._fused else F.relu(conv(h, edge_index))
            if self.use_batch_norm:
                h = self.bns[i](h)
            if self.use_layer_norm:
                h = self.lns[i](h)
            if (self.dropout > 0 and self.training and self.activation in (F.relu, torch.relu)
                    and hasattr(self.ops, "relu_dropout")):
                h = self.ops.relu_dropout(h, self.dropout, True)        # activation + dropout in one pass
            else:
                h = F.dropout(self.activation(h), p=self.dropout, training=self.training)
        h = self.conv_layers[-1](h, edge_index)
        return self.ops.scatter_mean(h, graph_of_node, dim=0)


class SCN(nn.Module):
    """Spectral-clustering net: GraphConv stack -> MLP logits -> MinCUT losses (model/hscn.py:19-64)."""

    def __init__(self, mp_units: Sequence[int], mp_act: str, num_features: int, num_clusters: int,
                 mlp_units: Sequence[int] = (), mlp_act: str = "identity",
                 ops: Optional[SimpleNamespace] = None):
        super().__init__()
        o = self.ops = ops or _default_ops()
        act = ACTIVATIONS[mp_act.lower()]
        sig = "x, edge_index, edge_weight -> x"
        dims = [num_features] + list(mp_units)
        steps: List = []
        for i in range(len(mp_units)):
            steps += [(o.GraphConv(dims[i], dims[i + 1]), sig), act]
        self.mp = o.Sequential("x, edge_index, edge_weight", steps)
        width = dims[-1]
        self.mlp = nn.Sequential()
        for units in mlp_units:  # the reference never advances `width` (Appendix B-9)
            self.mlp.append(o.Linear(width, units))
            self.mlp.append(ACTIVATIONS[mlp_act.lower()])
        self.mlp.append(o.Linear(width, num_clusters))
        # one GraphConv + activation + one Linear is what main.py:101-106 builds: operator sets that offer it run
        # that chain as one fused launch (same values within fp32 rounding; tests/test_gpu_parity.py)
        self._fusable = (len(mp_units) == 1 and len(mlp_units) == 0 and mp_act.lower() in ("elu", "relu", "tanh", "identity")
                         and hasattr(o, "scn_logits_fused"))
        self._act_name = mp_act.lower()
        self.fuse = os.environ.get("GHSCN_FUSED_SCN", "1") != "0"

    def logits(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor]) -> Tensor:
        """s = mlp(mp(x, edge_index, edge_weight))  (model/hscn.py:57-60)."""
        if self._fusable and self.fuse:
            s = self.ops.scn_logits_fused(x, edge_index, edge_weight, self.mp.module_0, self._act_name, self.mlp[0])
            if s is not None:
                return s
        return self.mlp(self.mp(x, edge_index, edge_weight))

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor]):
        """One graph per call, as train_clustering.py:44-47 drives it."""
        o = self.ops
        h = self.mp(x, edge_index, edge_weight)
        s = self.mlp(h)
        adj = o.to_dense_adj(edge_index)
        _, _, mc_loss, o_loss = o.dense_mincut_pool(h, adj, s)
        return torch.softmax(s, dim=-1), mc_loss, o_loss, adj

    def forward_batched(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor], batch: Tensor,
                        losses_tensor: bool = False, num_graphs: Optional[int] = None):
        """Throughput form: the whole `batch`/`ptr` mini-batch in one launch (losses are batch means).
        -> (s, mincut_loss, ortho_loss), or (s, losses[2]) with `losses_tensor` (CUDA operator set only).
        `num_graphs` restricts the losses to the first that many graphs (trailing padding graphs of a bucketed
        batch, train.GraphHSCNStep; CUDA operator set only)."""
        o = self.ops
        if self._fusable and self.fuse:
            s = self.logits(x, edge_index, edge_weight)
            h = x                       # the pooled features are not computed (want_out=False): only the shape matters
        else:
            h = self.mp(x, edge_index, edge_weight)
            s = self.mlp(h)
        kw = {} if num_graphs is None else {"num_graphs": num_graphs}
        if losses_tensor:
            _, _, both = o.mincut_pool_ragged(h, edge_index, s, batch, want_out=False, want_adj=False,
                                              losses_tensor=True, **kw)
            return s, both
        _, _, mc_loss, o_loss = o.mincut_pool_ragged(h, edge_index, s, batch, want_out=False, want_adj=False, **kw)
        return s, mc_loss, o_loss


def build_conv_relation(conv_type: str, hidden_channels: int, ops: SimpleNamespace):
    """model/hscn.py:117-125, including the case-sensitive 'GAT' test (Appendix B-10)."""
    dim = (-1, -1) if conv_type == "GAT" else -1
    cls = {"gcn": ops.GCNConv, "gat": ops.GATConv, "gin": ops.GINConv}[conv_type.lower()]
    return cls(dim, hidden_channels, add_self_loops=False, cached=False)


class HSCN(nn.Module):
    """Hetero GNN over local/virtual nodes (model/hscn.py:67-114)."""

    def __init__(self, lv_conv: str, ll_conv: str, vv_conv: str, activation: Callable, num_features: int,
                 hidden_channels: int, num_classes: int, num_layers: int,
                 ops: Optional[SimpleNamespace] = None):
        super().__init__()
        o = self.ops = ops or _default_ops()
        self.activation = activation
        self.convs = nn.ModuleList(
            o.HeteroConv({
                ("local", "to", "virtual"): build_conv_relation(lv_conv, hidden_channels, o),
                ("local", "to", "local"): build_conv_relation(ll_conv, hidden_channels, o),
                ("virtual", "to", "virtual"): build_conv_relation(vv_conv, hidden_channels, o),
            }, aggr="sum") for _ in range(num_layers))
        self.lin_1 = o.Linear(hidden_channels, hidden_channels)
        self.lin_2 = o.Linear(hidden_channels, num_classes)
        # "local" receives only the l->l GCN, so its `.relu()` can run in that layer's aggregation epilogue
        self._fused_local = self._fused_virtual = False
        self.first_layer_output: Optional[Tensor] = None
        if getattr(o, "fused_relu", False):
            lls = [c.convs["local__to__local"] for c in self.convs]
            if all(hasattr(c, "fuse_relu") for c in lls):
                for c in lls:
                    c.fuse_relu = True
                self._fused_local = True
            if all(hasattr(c, "fuse_relu_dst") for c in self.convs):      # "virtual": ReLU inside the HeteroConv
                for c in self.convs:
                    c.fuse_relu_dst = {"virtual"}
                self._fused_virtual = True

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict, batch) -> Tensor:
        return self.lin_2(self.forward_hidden(x_dict, edge_index_dict, batch))

    def forward_hidden(self, x_dict: Dict[str, Tensor], edge_index_dict, batch) -> Tensor:
        """model/hscn.py:102-111: everything in front of the output layer `lin_2` (the training step fuses that layer
        with the loss and their backward, ops.head_out_loss)."""
        # Operator sets whose HeteroConv runs the destination types on parallel CUDA streams (`last_streams`) keep
        # every type's tensors on its own stream across the layers: the "virtual" branch never feeds "local"
        # (model/hscn.py:84-94 has no virtual->local relation), so the caller's stream only waits for it once, at the
        # end -- or not at all when `defer_branch_join` is set and the training step joins after its optimizer step.
        streams = None
        main = None
        for conv in self.convs:
            if hasattr(conv, "defer_join"):
                conv.defer_join = True
            try:
                out = conv(x_dict, edge_index_dict)
            finally:
                if hasattr(conv, "defer_join"):
                    conv.defer_join = False
            streams = getattr(conv, "last_streams", None)
            if streams is not None and main is None:
                main = torch.cuda.current_stream()
            x_dict = {}
            for k, v in out.items():
                if (self._fused_local and k == "local") or (self._fused_virtual and k == "virtual"):
                    x_dict[k] = v
                elif streams is not None:
                    with torch.cuda.stream(streams[k]):
                        x_dict[k] = v.relu()
                else:
                    x_dict[k] = v.relu()
            if conv is self.convs[0]:
                # the activation every later parameter's gradient is complete behind (train.py overlaps the
                # data-parallel exchange of those gradients with the first layer's backward)
                self.first_layer_output = x_dict.get("local") if torch.is_grad_enabled() else None
        pooled = self.ops.global_mean_pool(x_dict["local"], batch["local"].batch)
        hidden = None
        if hasattr(self.ops, "linear_act"):          # bias + activation in the projection's epilogue (same values)
            hidden = self.ops.linear_act(self.lin_1, pooled, self.activation)
        if hidden is None:
            hidden = self.activation(self.lin_1(pooled))
        if streams is not None and not getattr(self, "defer_branch_join", False):
            for st in streams.values():
                if st is not main:
                    main.wait_stream(st)
        return hidden


def criterion(loss_fn: str, pred: Tensor, true: Tensor):
    """graph_hscn/loss.py:6-19 (incl. sigmoid score for L1, Appendix B-12)."""
    if loss_fn == "cross_entropy":
        if pred.ndim > 1 and true.ndim == 1:
            logp = F.log_softmax(pred, dim=-1)
            return F.nll_loss(logp, true), logp
        return F.binary_cross_entropy_with_logits(pred, true.float()), torch.sigmoid(pred)
    return F.l1_loss(pred, true), torch.sigmoid(pred)
