"""nn.Module restatement of the PyG layers the reference instantiates.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parameter names and shapes
follow PyG so state_dicts interchange with the product layers
(graph_hscn_b200.pyg.nn) and with a real PyG install:
  GCNConv:   lin.weight [out,in], bias [out]                 
  GraphConv: lin_rel.{weight,bias}, lin_root.weight          
  GATConv:   lin_src.weight, lin_dst.weight, att_src/att_dst [1,H,C], bias   
  HeteroConv: convs.<src__rel__dst>.*                        
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn.parameter import UninitializedParameter

from . import ops


def glorot(t: Tensor) -> None:
    if t is not None:
        stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        with torch.no_grad():
            t.uniform_(-stdv, stdv)


class Linear(nn.Module):
    """torch_geometric.nn.Linear (hscn.py:50-54,99-100): lazy in_channels=-1, y = x W^T + b."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        if in_channels > 0:
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        else:
            self.weight = UninitializedParameter()
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        if isinstance(self.weight, UninitializedParameter):
            return
        if self.weight_initializer == "glorot":
            glorot(self.weight)
        else:
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            if self.bias_initializer == "zeros":
                nn.init.zeros_(self.bias)
            else:
                bound = 1.0 / math.sqrt(self.weight.size(1)) if self.weight.size(1) > 0 else 0.0
                nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = x.size(-1)
            self.weight.materialize((self.out_channels, self.in_channels))
            self.reset_parameters()
        return F.linear(x, self.weight, self.bias)


class MessagePassing(nn.Module):
    """Type anchor only (config.py:9, mpnn.py:7 use it in annotations)."""


class GCNConv(MessagePassing):
    """SURVEY A.3 -- linear first, then normalised aggregation, then bias."""

    def __init__(self, in_channels: int, out_channels: int, improved: bool = False, cached: bool = False,
                 add_self_loops: bool = True, normalize: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = Linear(in_channels, out_channels, bias=False, weight_initializer="glorot")
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        if self.normalize:
            edge_index, edge_weight = ops.gcn_norm(edge_index, edge_weight, x.size(0), self.improved,
                                                   self.add_self_loops, "source_to_target", x.dtype)
        x = self.lin(x)
        out = ops.propagate_add(x, edge_index, edge_weight, x.size(0))
        if self.bias is not None:
            out = out + self.bias
        return out


class GraphConv(MessagePassing):
    """SURVEY A.4 -- aggregate at input width, then lin_rel(agg) + lin_root(x)."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, aggr: str = "add",
                 bias: bool = True, **kwargs):
        super().__init__()
        assert aggr == "add"
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_root = Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: Tensor, edge_weight: Optional[Tensor] = None, size=None) -> Tensor:
        if isinstance(x, Tensor):
            x = (x, x)
        n_dst = x[1].size(0) if x[1] is not None else (size[1] if size is not None else x[0].size(0))
        out = ops.propagate_add(x[0], edge_index, edge_weight, n_dst)
        out = self.lin_rel(out)
        if x[1] is not None:
            out = out + self.lin_root(x[1])
        return out


class GINConv(MessagePassing):
    """Listed in CONV_DICT (config.py:19-23); out = nn((1+eps) x_i + sum_j x_j)."""

    def __init__(self, nn_module: Callable, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.nn = nn_module
        self.initial_eps = eps
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([eps]))
        else:
            self.register_buffer("eps", torch.tensor([eps]))

    def forward(self, x, edge_index: Tensor, size=None) -> Tensor:
        if isinstance(x, Tensor) and x.dim() == 3:
            # PyG propagates along node_dim = -2: [K, N, C] (encoder/signnet.py:227-229) is aggregated per slice
            k, n, c = x.shape
            flat = x.permute(1, 0, 2).reshape(n, k * c)
            agg = ops.propagate_add(flat, edge_index, None, n).view(n, k, c).permute(1, 0, 2)
            return self.nn(agg + (1 + self.eps) * x)
        if isinstance(x, Tensor):
            x = (x, x)
        out = ops.propagate_add(x[0], edge_index, None, x[1].size(0))
        if x[1] is not None:
            out = out + (1 + self.eps) * x[1]
        return self.nn(out)


class GATConv(MessagePassing):
    """SURVEY A.8 (PyG 2.2/2.3 GATConv, edge_dim=None)."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, heads: int = 1,
                 concat: bool = True, negative_slope: float = 0.2, dropout: float = 0.0,
                 add_self_loops: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        if isinstance(in_channels, int):
            self.lin_src = Linear(in_channels, heads * out_channels, bias=False, weight_initializer="glorot")
            self.lin_dst = self.lin_src
        else:
            self.lin_src = Linear(in_channels[0], heads * out_channels, False, weight_initializer="glorot")
            self.lin_dst = Linear(in_channels[1], heads * out_channels, False, weight_initializer="glorot")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        glorot(self.att_src)
        glorot(self.att_dst)
        if bias and concat:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        elif bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index: Tensor, edge_attr=None, size=None) -> Tensor:
        H, C = self.heads, self.out_channels
        if isinstance(x, Tensor):
            x_src = x_dst = self.lin_src(x).view(-1, H, C)
        else:
            x_src, x_dst = x
            x_src = self.lin_src(x_src).view(-1, H, C)
            if x_dst is not None:
                x_dst = self.lin_dst(x_dst).view(-1, H, C)
        alpha_src = (x_src * self.att_src).sum(dim=-1)
        alpha_dst = None if x_dst is None else (x_dst * self.att_dst).sum(-1)
        if self.add_self_loops:
            num_nodes = x_src.size(0)
            if x_dst is not None:
                num_nodes = min(num_nodes, x_dst.size(0))
            mask = edge_index[0] != edge_index[1]
            loops = torch.arange(num_nodes, dtype=torch.long).unsqueeze(0).repeat(2, 1)
            edge_index = torch.cat([edge_index[:, mask], loops], dim=1)
        row, col = edge_index[0], edge_index[1]
        n_dst = x_dst.size(0) if x_dst is not None else (size[1] if size is not None else x_src.size(0))
        alpha = alpha_src.index_select(0, row)
        if alpha_dst is not None:
            alpha = alpha + alpha_dst.index_select(0, col)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = ops.segment_softmax(alpha, col, n_dst)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = alpha.unsqueeze(-1) * x_src.index_select(0, row)
        out = ops.scatter_sum(msg, col, dim=0, dim_size=n_dst)
        out = out.view(-1, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out


class HeteroConv(nn.Module):
    """iterate edge_index_dict in order, group outputs per destination type."""

    def __init__(self, convs: Dict[Tuple[str, str, str], nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict: Dict[Tuple[str, str, str], Tensor]
                ) -> Dict[str, Tensor]:
        out_dict: Dict[str, List[Tensor]] = defaultdict(list)
        for edge_type, edge_index in edge_index_dict.items():
            src, rel, dst = edge_type
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            conv = self.convs[key]
            if src == dst:
                out = conv(x_dict[src], edge_index)
            else:
                out = conv((x_dict[src], x_dict[dst]), edge_index)
            out_dict[dst].append(out)
        result = {}
        for key, xs in out_dict.items():
            stacked = torch.stack(xs, dim=0)
            agg = "sum" if self.aggr in ("sum", "add") else self.aggr
            red = getattr(torch, agg)(stacked, dim=0)
            result[key] = red[0] if isinstance(red, tuple) else red
        return result


class Sequential(nn.Module):
    """torch_geometric.nn.Sequential as used at hscn.py:30-45 ('x, edge_index, edge_weight' signature)."""

    def __init__(self, input_args: str, modules: List[Union[Tuple[Callable, str], Callable]]):
        super().__init__()
        self._inputs = [a.strip() for a in input_args.split(",")]
        self._steps: List[Tuple[str, List[str], List[str]]] = []
        prev_out = [self._inputs[0]]
        for i, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                fn, desc = entry
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",")]
                outs = [a.strip() for a in rhs.split(",")]
            else:
                fn, ins, outs = entry, list(prev_out), list(prev_out)
            name = f"module_{i}"
            if isinstance(fn, nn.Module):
                self.add_module(name, fn)
            else:
                object.__setattr__(self, name, fn)
            self._steps.append((name, ins, outs))
            prev_out = outs

    def forward(self, *args):
        env = dict(zip(self._inputs, args))
        out = None
        for name, ins, outs in self._steps:
            out = getattr(self, name)(*[env[k] for k in ins])
            if len(outs) == 1:
                env[outs[0]] = out
            else:
                for k, v in zip(outs, out):
                    env[k] = v
        return out
