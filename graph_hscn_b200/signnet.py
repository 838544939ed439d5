"""Host-side mirror of the reference's SignNet positional-encoding encoder, parameterised by an
operator namespace like `models.py`.

Reference: graph_hscn/encoder/signnet.py -- MLP :11-85, GIN :88-162, GINDeepSigns :165-214, MaskedGINDeepSigns
:217-287, SignNetNodeEncoder :290-381.  Same constructor arguments and submodule / parameter names (state_dicts
interchange), same forward values; tests/golden/posenc.pt pins mirror == the reference's source text.

Two things differ from the reference on purpose and are recorded here:
  * `MLP.__init__` looks the activation up as `ACT_DICT["activation"]` (signnet.py:49), a KeyError for every
    argument -- the encoder cannot be constructed as shipped.  The mirror uses the evident
    intent `ACT_DICT[activation]`; the golden fixture runs the reference with that one dictionary key added.
  * `MaskedGINDeepSigns` builds its mask with one Python list entry per node and per graph (signnet.py:250-275); here
    it is a broadcast compare against the per-node graph sizes from `ptr` (same values, no host loop).
The three-dimensional [K, N, C] stack goes through `GINConv` as ONE aggregation over [N, K*C] (pyg/nn.py).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .models import ACTIVATIONS, _default_ops


def _batch_norm(bn: nn.BatchNorm1d, x: Tensor) -> Tensor:
    """signnet.py:61-67 / :151-158: BatchNorm1d over the channel dim of [N, C] or [K, N, C]."""
    if x.ndim == 2:
        return bn(x)
    if x.ndim == 3:
        return bn(x.transpose(2, 1)).transpose(2, 1)
    raise ValueError("Invalid dimension of x")


class MLP(nn.Module):
    """signnet.py:11-85."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int,
                 use_bn: bool = False, use_ln: bool = False, dropout: float = 0.5, activation: str = "relu",
                 residual: bool = False):
        super().__init__()
        widths = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.fcs = nn.ModuleList(nn.Linear(widths[i], widths[i + 1]) for i in range(num_layers))
        if use_bn:
            self.bns = nn.ModuleList(nn.BatchNorm1d(hidden_channels) for _ in range(num_layers - 1))
        if use_ln:
            self.lns = nn.ModuleList(nn.LayerNorm(hidden_channels) for _ in range(num_layers - 1))
        self.activation = ACTIVATIONS[activation]
        self.dropout, self.use_bn, self.use_ln, self.residual = dropout, use_bn, use_ln, residual

    def forward(self, x: Tensor) -> Tensor:
        prev = x
        for i in range(len(self.fcs) - 1):
            x = self.activation(self.fcs[i](x))
            if self.use_bn:
                x = _batch_norm(self.bns[i], x)
            if self.use_ln:
                x = self.lns[i](x)
            if self.residual and prev.shape == x.shape:
                x = x + prev
            x = F.dropout(x, p=self.dropout, training=self.training)
            prev = x
        x = self.fcs[-1](x)
        if self.residual and prev.shape == x.shape:
            x = x + prev
        return x


class GIN(nn.Module):
    """signnet.py:88-162: GINConv(MLP) layers; the first and the last always exist, so `n_layers` <= 2 gives two."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, n_layers: int, use_bn: bool = True,
                 dropout: float = 0.5, activation: str = "relu", ops: Optional[SimpleNamespace] = None):
        super().__init__()
        ops = ops or _default_ops()
        kw = dict(use_bn=use_bn, dropout=dropout, activation=activation)
        specs = ([(in_channels, hidden_channels, hidden_channels, 1)]
                 + [(hidden_channels, hidden_channels, hidden_channels, 1)] * max(n_layers - 2, 0)
                 + [(hidden_channels, hidden_channels, out_channels, 2)])
        self.layers = nn.ModuleList(ops.GINConv(MLP(i, h, o, depth, **kw)) for i, h, o, depth in specs)
        if use_bn:
            self.bns = nn.ModuleList(nn.BatchNorm1d(hidden_channels) for _ in range(len(specs) - 1))
        self.use_bn, self.dropout = use_bn, dropout

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        for i, layer in enumerate(self.layers):
            if i != 0:
                x = F.dropout(x, p=self.dropout, training=self.training)
                if self.use_bn:
                    x = _batch_norm(self.bns[i - 1], x)
            x = layer(x, edge_index)
        return x


class GINDeepSigns(nn.Module):
    """signnet.py:165-214: rho(concat_k [phi(v_k) + phi(-v_k)])."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int, k: int,
                 dim_pe: int, rho_num_layers: int, use_bn: bool = False, use_ln: bool = False, dropout: float = 0.5,
                 activation: str = "relu", ops: Optional[SimpleNamespace] = None):
        super().__init__()
        self.enc = GIN(in_channels, hidden_channels, out_channels, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation, ops=ops)
        self.rho = MLP(out_channels * k, hidden_channels, dim_pe, rho_num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)

    def forward(self, x: Tensor, edge_index: Tensor, batch_index: Tensor) -> Tensor:
        n = x.shape[0]
        x = x.transpose(0, 1)                                       # [N, K, in] -> [K, N, in]
        x = self.enc(x, edge_index) + self.enc(-x, edge_index)
        return self.rho(x.transpose(0, 1).reshape(n, -1))


class MaskedGINDeepSigns(nn.Module):
    """signnet.py:217-287: rho(sum_{k < n_graph} [phi(v_k) + phi(-v_k)]) -- frequencies past the graph's node count
    (the NaN padding of posenc.py:66-76) are masked out."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int, dim_pe: int,
                 rho_num_layers: int, use_bn: bool = False, use_ln: bool = False, dropout: float = 0.5,
                 activation: str = "relu", ops: Optional[SimpleNamespace] = None):
        super().__init__()
        self.enc = GIN(in_channels, hidden_channels, out_channels, num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation, ops=ops)
        self.rho = MLP(out_channels, hidden_channels, dim_pe, rho_num_layers, use_bn=use_bn, dropout=dropout,
                       activation=activation)

    @staticmethod
    def batched_n_nodes(batch_index: Tensor, num_graphs: Optional[int] = None) -> Tensor:
        """signnet.py:250-260: for every node, the node count of its graph (no host loop over graphs)."""
        if num_graphs is None:
            num_graphs = int(batch_index.max()) + 1
        counts = torch.zeros(num_graphs, dtype=torch.long, device=batch_index.device)
        counts.scatter_add_(0, batch_index, torch.ones_like(batch_index))
        return counts[batch_index]

    def forward(self, x: Tensor, edge_index: Tensor, batch_index: Tensor, num_graphs: Optional[int] = None) -> Tensor:
        k = x.shape[1]
        x = x.transpose(0, 1)
        x = self.enc(x, edge_index) + self.enc(-x, edge_index)      # [K, N, out]
        x = x.transpose(0, 1)                                       # [N, K, out]
        keep = torch.arange(k, device=x.device).unsqueeze(0) < self.batched_n_nodes(batch_index, num_graphs).unsqueeze(1)
        x = (x * keep.unsqueeze(-1)).sum(dim=1)
        return self.rho(x)


class SignNetNodeEncoder(nn.Module):
    """signnet.py:290-381.  `cfg` carries the PEConfig fields (config/config.py:115-130): dim_pe, model, layers,
    post_layers, eigen_max_freqs, phi_hidden_dim, phi_out_dim, pass_as_var, use_bn."""

    def __init__(self, cfg, dim_in: int, dim_emb: int, expand_x: bool = True, ops: Optional[SimpleNamespace] = None):
        super().__init__()
        if cfg.model not in ("MLP", "DeepSet"):
            raise ValueError(f"Unexpected SignNet model {cfg.model}")
        if cfg.post_layers < 1:
            raise ValueError("Num layers in rho model has to be positive.")
        if dim_emb - cfg.dim_pe < 1:
            raise ValueError(f"SignNet PE size {cfg.dim_pe} is too large for desired embedding size of {dim_emb}.")
        self.model_type, self.pass_as_var, self.expand_x = cfg.model, cfg.pass_as_var, expand_x
        if expand_x:
            self.linear_x = nn.Linear(dim_in, dim_emb - cfg.dim_pe)
        common = dict(in_channels=1, hidden_channels=cfg.phi_hidden_dim, out_channels=cfg.phi_out_dim,
                      num_layers=cfg.layers, dim_pe=cfg.dim_pe, rho_num_layers=cfg.post_layers, use_bn=cfg.use_bn,
                      dropout=0.0, activation="relu", ops=ops)
        if cfg.model == "MLP":
            self.sign_inv_net = GINDeepSigns(k=cfg.eigen_max_freqs, **common)
        else:
            self.sign_inv_net = MaskedGINDeepSigns(**common)

    def forward(self, batch):
        if not (hasattr(batch, "eigvals_sn") and hasattr(batch, "eigvecs_sn")):
            raise ValueError(f"Precomputed eigen values and vectors are required for {self.__class__.__name__}; "
                             "set config 'posenc_SignNet.enable' to True")
        pos_enc = torch.nan_to_num(batch.eigvecs_sn.unsqueeze(-1), nan=0.0, posinf=float("inf"),
                                   neginf=float("-inf"))           # signnet.py:358-360: NaN padding -> 0
        pos_enc = self.sign_inv_net(pos_enc, batch.edge_index, batch.batch)
        h = self.linear_x(batch.x.to(torch.float32)) if self.expand_x else batch.x
        batch.x = torch.cat((h, pos_enc), 1)
        if self.pass_as_var:
            batch.pe_SignNet = pos_enc
        return batch
