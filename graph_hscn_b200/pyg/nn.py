"""PyG-named layers backed by the sm_100a kernels (drop-in boundary).

Constructor signatures, parameter names/shapes and forward semantics mirror torch_geometric so the
reference's model code (model/mpnn.py, model/hscn.py) instantiates and calls them unchanged and
state_dicts interchange with a real PyG install:
  GCNConv   lin.weight [out,in], bias [out]                        model/mpnn.py:29-32,52,59; hscn.py:88-93
  GraphConv lin_rel.{weight,bias}, lin_root.weight                 model/hscn.py:32-34,40-41
  GATConv   lin_src/lin_dst.weight, att_src/att_dst [1,H,C], bias  model/hscn.py:85-87
  HeteroConv convs.<src__rel__dst>.*                               model/hscn.py:83-96,109
Dense projections (x @ W^T) go through gemm.linear: hand-written tcgen05 3xTF32 kernels for the h x h layers,
streaming kernels for the tall-skinny ones, fp32 library GEMM otherwise (TF32 off: parity is 1e-5).
CPU tensors / CPU-resident parameters are staged to the GPU when auto-device staging is on (pyg/_device.py:
the reference's MPNN path, train/train.py:78-81, never moves its model or batches); otherwise they raise.
"""
from __future__ import annotations

import contextlib
import os

import math
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn.parameter import UninitializedParameter

from .. import gemm, ops
from ..structure import _require_cuda, structure_cache
from . import _device


def glorot(t: Optional[Tensor]) -> None:
    if t is not None:
        stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        with torch.no_grad():
            t.uniform_(-stdv, stdv)


class Linear(nn.Module):
    """torch_geometric.nn.Linear: supports lazy `in_channels=-1` (hscn.py:50-54,99-100)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight_initializer, self.bias_initializer = weight_initializer, bias_initializer
        self.weight = (nn.Parameter(torch.empty(out_channels, in_channels)) if in_channels > 0
                       else UninitializedParameter())
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

    def materialize(self, in_channels: int, device) -> None:
        """Resolve a lazy `in_channels=-1` weight (PyG does this in a forward pre-hook)."""
        if isinstance(self.weight, UninitializedParameter):
            self.in_channels = in_channels
            self.weight.materialize((self.out_channels, self.in_channels), device=device)
            self.reset_parameters()

    def forward(self, x: Tensor) -> Tensor:
        self.materialize(x.size(-1), x.device)
        dev, back = _device.plan(x, self.weight)
        if dev is not None:
            y = gemm.linear(_device.to_dev(x, dev), _device.to_dev(self.weight, dev), _device.to_dev(self.bias, dev))
            return _device.back_to(y, back)
        return gemm.linear(x, self.weight, self.bias)

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, bias={self.bias is not None}"


class MessagePassing(nn.Module):
    """Base type (config/config.py:9 and model/mpnn.py:7 annotate with it)."""


def _aggregate(x_src: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor], num_dst: int,
               normalize: bool = False, improved: bool = False, add_self_loops: bool = False,
               bias: Optional[Tensor] = None, relu: bool = False) -> Tensor:
    """out[i] = act(sum_{e: col_e = i} w_e * x_src[row_e] + bias): CSR lookup + K2 SpMM (ReLU in its epilogue)."""
    _require_cuda(x_src, edge_index)
    st = structure_cache().graph(edge_index, x_src.size(0), num_dst, add_self_loops)
    w_param = None
    if edge_weight is not None and edge_weight.requires_grad and not normalize:
        # differentiable explicit weights: gather to slot order with autograd, unit loops not involved
        w_param = edge_weight.float()[st.by_dst.perm.long()]
        w, w_t = w_param, edge_weight.detach().float()[st.by_src.perm.long()]
    else:
        if edge_weight is not None and edge_weight.requires_grad:
            raise NotImplementedError("gradients w.r.t. edge_weight through gcn_norm are not on the reference's path")
        w, w_t, _ = st.weights(edge_weight, normalize=normalize, improved=improved)
    d, s = st.by_dst, st.by_src
    if x_src.dtype != torch.float32:
        x_src = x_src.float()
    return torch.ops.ghscn.spmm(d.rowptr, d.col, w, s.rowptr, s.col, w_t, x_src, bias, relu)


class GCNConv(MessagePassing):
    """out = D^-1/2 (A [+ I]) D^-1/2 (x W^T) + b  (SURVEY A.3).  gcn_norm is folded into the batch CSR,
    which is built once per `edge_index` tensor and shared by all layers and by the backward."""

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
        # extension (not in PyG): a caller that applies ReLU right after this layer may set this flag and skip
        # its own ReLU; the activation then runs in the aggregation kernel's epilogue (same values)
        self.fuse_relu = False

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        dev, back = _device.plan(x, edge_index, edge_weight, self.bias)
        bias = self.bias
        if dev is not None:
            x, edge_weight, bias = _device.to_dev(x, dev), _device.to_dev(edge_weight, dev), _device.to_dev(bias, dev)
            edge_index = _device.index_to_dev(edge_index, dev)
        h = self.lin(x)
        out = _aggregate(h, edge_index, edge_weight, x.size(0), normalize=self.normalize,
                         improved=self.improved, add_self_loops=self.add_self_loops and self.normalize,
                         bias=bias, relu=self.fuse_relu)
        return _device.back_to(out, back)


class GraphConv(MessagePassing):
    """out = lin_rel(sum_j w_ji x_j) + lin_root(x_i)  (SURVEY A.4): aggregate at input width first."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, aggr: str = "add",
                 bias: bool = True, **kwargs):
        super().__init__()
        if aggr != "add":
            raise NotImplementedError("GraphConv aggr must be 'add' (the reference's default)")
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = Linear(in_channels[0], out_channels, bias=bias)
        self.lin_root = Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, edge_index: Tensor, edge_weight: Optional[Tensor] = None, size=None) -> Tensor:
        if isinstance(x, Tensor):
            x = (x, x)
        dev, back = _device.plan(x[0], x[1], edge_index, edge_weight)
        if dev is not None:
            x = (_device.to_dev(x[0], dev), _device.to_dev(x[1], dev))
            edge_index, edge_weight = _device.index_to_dev(edge_index, dev), _device.to_dev(edge_weight, dev)
        n_dst = x[1].size(0) if x[1] is not None else (size[1] if size is not None else x[0].size(0))
        out = self.lin_rel(_aggregate(x[0], edge_index, edge_weight, n_dst))
        if x[1] is not None:
            out = out + self.lin_root(x[1])
        return _device.back_to(out, back)


class GINConv(MessagePassing):
    """Listed in CONV_DICT (config/config.py:19-23): out = nn((1 + eps) x_i + sum_j x_j)."""

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
            # PyG propagates along node_dim = -2: the [K, N, C] stack of encoder/signnet.py:227-229 is ONE
            # aggregation over [N, K*C] (the sum over neighbours is per column), then `nn` on the last dim
            k, n, c = x.shape
            dev, back = _device.plan(x, edge_index)
            flat = _device.to_dev(x, dev).permute(1, 0, 2).reshape(n, k * c)
            agg = _aggregate(flat, _device.index_to_dev(edge_index, dev), None, n).view(n, k, c).permute(1, 0, 2)
            agg = _device.back_to(agg, back)
            return self.nn(agg + (1 + _device.to_dev(self.eps, agg.device)) * x)
        if isinstance(x, Tensor):
            x = (x, x)
        dev, back = _device.plan(x[0], x[1], edge_index)
        if dev is not None:
            x = (_device.to_dev(x[0], dev), _device.to_dev(x[1], dev))
            edge_index = _device.index_to_dev(edge_index, dev)
        out = _aggregate(x[0], edge_index, None, x[1].size(0))
        if x[1] is not None:
            out = out + (1 + _device.to_dev(self.eps, out.device)) * x[1]
        return _device.back_to(self.nn(out), back)


class GATConv(MessagePassing):
    """Graph attention.  The bipartite single-head form GATConv((-1,-1), C, add_self_loops=False)
    on ("local","to","virtual") is the reference's cluster pool; heads > 1 run head by head on the same kernels."""

    def __init__(self, in_channels: Union[int, Tuple[int, int]], out_channels: int, heads: int = 1,
                 concat: bool = True, negative_slope: float = 0.2, dropout: float = 0.0,
                 add_self_loops: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("attention dropout is not on the reference's path")
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
        if bias:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)

    def forward(self, x, edge_index: Tensor, edge_attr=None, size=None) -> Tensor:
        if edge_attr is not None:
            raise NotImplementedError("GATConv edge features are not on the reference's path")
        if isinstance(x, Tensor):
            x_src = x_dst = x
        else:
            x_src, x_dst = x
        dev, back = _device.plan(x_src, x_dst, edge_index, self.att_src)
        if dev is not None:
            return _device.back_to(self._forward_staged(x, x_src, x_dst, edge_index, size, dev), back)
        return self._forward_cuda(x_src, x_dst, edge_index, size, None)

    def _params(self, dev=None):
        ls, ld = self.lin_src.weight, self.lin_dst.weight
        p = dict(w_src=ls, w_dst=ld, att_src=self.att_src, att_dst=self.att_dst, bias=self.bias)
        return p if dev is None else {k: _device.to_dev(v, dev) for k, v in p.items()}

    def _forward_staged(self, x, x_src, x_dst, edge_index, size, dev):
        self.lin_src.materialize(x_src.size(-1), x_src.device)
        if x_dst is not None:
            self.lin_dst.materialize(x_dst.size(-1), x_dst.device)
        xs = _device.to_dev(x_src, dev)
        xd = xs if x_dst is x_src else _device.to_dev(x_dst, dev)
        return self._forward_cuda(xs, xd, _device.index_to_dev(edge_index, dev), size, self._params(dev))

    def _forward_cuda(self, x_src, x_dst, edge_index, size, prm) -> Tensor:
        _require_cuda(x_src, edge_index)
        if x_src.dtype != torch.float32:
            x_src = x_src.float()
        self.lin_src.materialize(x_src.size(-1), x_src.device)
        if x_dst is not None:
            if x_dst.dtype != torch.float32:
                x_dst = x_dst.float()
            self.lin_dst.materialize(x_dst.size(-1), x_dst.device)
        if prm is None:
            prm = self._params()
        w_src, w_dst, att_src, att_dst, bias = (prm[k] for k in ("w_src", "w_dst", "att_src", "att_dst", "bias"))
        n_src = x_src.size(0)
        n_dst = x_dst.size(0) if x_dst is not None else (size[1] if size is not None else n_src)
        if self.add_self_loops and n_src != n_dst:
            raise NotImplementedError("add_self_loops on a bipartite relation is not on the reference's path")
        st = structure_cache().graph(edge_index, n_src, n_dst, self.add_self_loops)
        d, s = st.by_dst, st.by_src
        # attention scores from x (W^T att) and the pooled sum at the input width: no [N,F]x[F,H] projection
        if self.heads == 1:
            return ops.GatPoolInputWidth.apply(
                x_src, x_dst, w_src, w_dst if x_dst is not None else None,
                att_src.view(-1), att_dst.view(-1), bias, float(self.negative_slope),
                d.rowptr, d.col, s.rowptr, s.col, st.slot_map_t)
        # heads > 1 (never built by the reference's configs)
        C = self.out_channels

        def head_by_head(xs_, xd_):
            """every head is the single-head operator on its own row block of lin_src / lin_dst and its own attention
            vectors, over the same structure (differentiable; the fused forward's backward runs this)"""
            parts = []
            for h in range(self.heads):
                rows = slice(h * C, (h + 1) * C)
                parts.append(ops.GatPoolInputWidth.apply(
                    xs_, xd_, w_src[rows], w_dst[rows] if xd_ is not None else None,
                    att_src[0, h], att_dst[0, h], None, float(self.negative_slope),
                    d.rowptr, d.col, s.rowptr, s.col, st.slot_map_t))
            return torch.cat(parts, dim=1)

        f_in = x_src.size(1)
        from .._lib import lib as _lib
        if (FUSED_MULTIHEAD and x_src.is_cuda and (x_dst is None or x_dst.size(1) == f_in)
                and _lib().query("ghscn_gat_pool_fused_supported", f_in, x_src.stride(0), f_in)):
            shared = x_dst is x_src
            meta = dict(heads=self.heads, out_channels=C, num_dst=n_dst, by_dst=d, slope=self.negative_slope,
                        unfused=head_by_head, shared_x=shared,
                        params=(w_src, None if (x_dst is None or w_dst is w_src) else w_dst, att_src, att_dst))
            cat = ops.GatMultiHead.apply(x_src, x_dst, w_src, w_dst if x_dst is not None else None, att_src, att_dst,
                                         meta)
            out = cat if self.concat else cat.view(n_dst, self.heads, C).mean(dim=1)
            return out + bias if bias is not None else out
        outs = []
        for h in range(self.heads):
            rows = slice(h * C, (h + 1) * C)
            outs.append(ops.GatPoolInputWidth.apply(
                x_src, x_dst, w_src[rows], w_dst[rows] if x_dst is not None else None,
                att_src[0, h], att_dst[0, h], None, float(self.negative_slope),
                d.rowptr, d.col, s.rowptr, s.col, st.slot_map_t))
        out = torch.cat(outs, dim=1) if self.concat else torch.stack(outs, dim=1).mean(dim=1)
        return out + bias if bias is not None else out


FUSED_VIRTUAL = os.environ.get("GHSCN_FUSED_VIRTUAL", "1") != "0"
FUSED_MULTIHEAD = os.environ.get("GHSCN_FUSED_MULTIHEAD", "1") != "0"
PARALLEL_BRANCHES = os.environ.get("GHSCN_PARALLEL_BRANCHES", "1") != "0"
PARALLEL_RELATIONS = os.environ.get("GHSCN_PARALLEL_RELATIONS", "1") != "0"
_SIDE_STREAMS: Dict[Tuple[str, int], List["torch.cuda.Stream"]] = {}


def branch_stream(device: torch.device, index: int = 0) -> "torch.cuda.Stream":
    """The side stream HeteroConv uses for its (index + 2)-th destination type on `device`."""
    return _side_streams(device, index + 1)[index]


_FORKED: Dict[Tuple[str, int], List["torch.cuda.Stream"]] = {}


def _forked(device: torch.device, stream: "torch.cuda.Stream") -> None:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    lst = _FORKED.setdefault(key, [])
    if all(s is not stream for s in lst):
        lst.append(stream)


def take_forked_streams(device: torch.device) -> List["torch.cuda.Stream"]:
    """Side streams HeteroConv has forked on `device` since the last call (and forgets them).  A caller that defers
    the joins (train.GraphHSCNStep) waits for exactly these at the end of its step -- inside a CUDA-graph capture
    only streams that took part in the capture may be waited for."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    return _FORKED.pop(key, [])


def _uses(t, stream) -> None:
    """Tells the caching allocator that `t` is read on `stream` (a no-op for the stream that allocated it): the block
    is not recycled under the reader, eagerly (event-deferred free) or inside a capture (free deferred to its end)."""
    if isinstance(t, Tensor) and t.is_cuda:
        t.record_stream(stream)


# priority of the branch streams (-1 = high): inherited by the kernel nodes of a captured graph.  Measured: high priority
# for the side chain 0.652 ms/step vs 0.638 at equal priority; confining the side chain to a CUDA green context of
# 32 / 48 SMs (torch.cuda.GreenContext) 0.903 / 0.778 ms -- its kernels need most of the GPU for a short time each.
SIDE_STREAM_PRIORITY = int(os.environ.get("GHSCN_SIDE_PRIORITY", "0"))


def _side_streams(device: torch.device, n: int):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    pool = _SIDE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device, priority=SIDE_STREAM_PRIORITY))
    return pool[:n]


class HeteroConv(nn.Module):
    """Runs one conv per relation in `edge_index_dict` order and sums per destination type (A.9)."""

    def __init__(self, convs: Dict[Tuple[str, str, str], nn.Module], aggr: Optional[str] = "sum"):
        super().__init__()
        self.convs = nn.ModuleDict({"__".join(k): v for k, v in convs.items()})
        self.aggr = aggr
        # set by a caller that keeps each destination type's tensors on its branch stream and joins later
        # (models.HSCN / train.GraphHSCNStep); `last_streams` tells it which stream owns which output
        self.defer_join = False
        self.last_streams: Optional[Dict[str, "torch.cuda.Stream"]] = None
        # extension (not in PyG): destination types whose output the caller passes through ReLU right away; the
        # activation then runs inside this layer (in the fused operator's epilogue where there is one; same values)
        self.fuse_relu_dst: set = set()

    def _fused_virtual_plan(self, x_dict, edge_index_dict):
        """Destination types that receive exactly {GCNConv on dst->dst (normalised, no self loops), single-head bipartite
        GATConv on src->dst (no self loops)} under aggr='sum' -- the reference's "virtual" type (model/hscn.py:83-96)
        -- are computed by ONE fused operator (ops.VirtualLayerFused).  -> {dst: (vv_type, lv_type)}"""
        if not FUSED_VIRTUAL or self.aggr not in ("sum", "add"):
            return {}
        by_dst: Dict[str, List[Tuple[str, str, str]]] = defaultdict(list)
        for et in edge_index_dict:
            if "__".join(et) in self.convs:
                by_dst[et[2]].append(et)
        plan = {}
        lib_ = None
        for dst, ets in by_dst.items():
            if len(ets) != 2:
                continue
            same = [e for e in ets if e[0] == dst]
            cross = [e for e in ets if e[0] != dst]
            if len(same) != 1 or len(cross) != 1 or ets[0] != same[0]:      # sum order: dst->dst first, then src->dst
                continue
            gcn, gat = self.convs["__".join(same[0])], self.convs["__".join(cross[0])]
            if not (isinstance(gcn, GCNConv) and isinstance(gat, GATConv)):
                continue
            if not (gcn.normalize and not gcn.add_self_loops and not gcn.improved and not gcn.fuse_relu
                    and gat.heads == 1 and not gat.add_self_loops and gat.lin_src is not gat.lin_dst):
                continue
            xs, xd = x_dict.get(cross[0][0]), x_dict.get(dst)
            if not (isinstance(xs, Tensor) and isinstance(xd, Tensor) and xs.is_cuda and xd.is_cuda
                    and xs.dtype == torch.float32 and xd.dtype == torch.float32 and xs.dim() == 2 and xd.dim() == 2
                    and xs.size(1) == xd.size(1) and xd.size(0) > 0):
                continue
            if lib_ is None:
                from .._lib import lib as _lib
                lib_ = _lib()
            f = xs.size(1)
            if not lib_.query("ghscn_gat_pool_fused_supported", f, xs.stride(0), f):
                continue
            plan[dst] = (same[0], cross[0])
        return plan

    def _fused_virtual(self, dst: str, vv_type, lv_type, x_dict, edge_index_dict, relu: bool) -> Tensor:
        gcn, gat = self.convs["__".join(vv_type)], self.convs["__".join(lv_type)]
        xs, xd = x_dict[lv_type[0]], x_dict[dst]
        vv_index, lv_index = edge_index_dict[vv_type], edge_index_dict[lv_type]
        gcn.lin.materialize(xd.size(-1), xd.device)                 # lazy parameters in PyG's (relation) order
        gat.lin_src.materialize(xs.size(-1), xs.device)
        gat.lin_dst.materialize(xd.size(-1), xd.device)
        V = xd.size(0)
        st_vv = structure_cache().graph(vv_index, V, V, False)
        vv_w, _, _ = st_vv.weights(None, normalize=True, need_transpose=False)
        st_lv = structure_cache().graph(lv_index, xs.size(0), V, False)
        meta = dict(gat=gat, gcn=gcn, vv_index=vv_index, lv_index=lv_index, lv_by_dst=st_lv.by_dst,
                    vv_by_dst=st_vv.by_dst, vv_w=vv_w, slope=gat.negative_slope, relu=bool(relu))
        return ops.VirtualLayerFused.apply(xs, xd, gat.lin_src.weight, gat.lin_dst.weight, gat.att_src.view(-1),
                                           gat.att_dst.view(-1), gat.bias, gcn.lin.weight, gcn.bias, meta)

    def forward(self, x_dict: Dict[str, Tensor], edge_index_dict: Dict[Tuple[str, str, str], Tensor]
                ) -> Dict[str, Tensor]:
        fused = self._fused_virtual_plan(x_dict, edge_index_dict)
        fused_types = {et for pair in fused.values() for et in pair}
        outs: Dict[str, List[Tensor]] = defaultdict(list)
        dsts: List[str] = []
        for edge_type in edge_index_dict:
            if "__".join(edge_type) in self.convs and edge_type[2] not in dsts:
                dsts.append(edge_type[2])
        # The relations of different destination types are independent (l->l feeds "local"; v->v and l->v feed
        # "virtual"): each destination type gets its own CUDA stream, forked from and joined back to the caller's
        # stream, so the small virtual-node kernels overlap the large local ones (also inside a captured CUDA
        # graph, where the fork/join become graph dependencies).  The relations are still visited in
        # `edge_index_dict` order (lazy parameters initialise in PyG's order) and run the same kernels: results
        # are bit-identical to the serial schedule (scripts/parallel_branch_check.py).
        streams = None
        if PARALLEL_BRANCHES and len(dsts) > 1 and all(t.is_cuda for t in x_dict.values()):
            main = torch.cuda.current_stream()
            sides = _side_streams(main.device, len(dsts) - 1)
            streams = {dsts[0]: main}
            for i, dst in enumerate(dsts[1:]):
                sides[i].wait_stream(main)
                _forked(main.device, sides[i])
                streams[dst] = sides[i]
        # Relations that share a destination type (v->v and l->v) are independent up to the final sum: the first one
        # runs on the destination's stream, every further one on a stream of its own that waits for the caller's
        # stream and for the destination's stream as they stand now (its inputs), and is joined before the sum.
        rel_streams: Dict[Tuple[str, str, str], "torch.cuda.Stream"] = {}
        joins: Dict[str, List["torch.cuda.Stream"]] = defaultdict(list)
        if streams is not None:
            extra = len(dsts) - 1
            for edge_type in edge_index_dict:
                if "__".join(edge_type) not in self.convs:
                    continue
                dst = edge_type[2]
                if not PARALLEL_RELATIONS or edge_type in fused_types or all(t[2] != dst for t in rel_streams):
                    rel_streams[edge_type] = streams[dst]
                    continue
                r = _side_streams(main.device, extra + 1)[extra]
                extra += 1
                r.wait_stream(main)
                _forked(main.device, r)
                if streams[dst] is not main:
                    r.wait_stream(streams[dst])
                rel_streams[edge_type] = r
                joins[dst].append(r)
        relu_done = set()
        for edge_type, edge_index in edge_index_dict.items():
            src, _, dst = edge_type
            key = "__".join(edge_type)
            if key not in self.convs:
                continue
            if edge_type in fused_types:
                vv_type, lv_type = fused[dst]
                if edge_type != lv_type:                 # run once, at the second (src -> dst) relation of the pair
                    continue
                if streams is not None:
                    _uses(x_dict[src], streams[dst])
                    _uses(x_dict[dst], streams[dst])
                with torch.cuda.stream(streams[dst]) if streams is not None else contextlib.nullcontext():
                    out = self._fused_virtual(dst, vv_type, lv_type, x_dict, edge_index_dict,
                                              relu=dst in self.fuse_relu_dst)
                if dst in self.fuse_relu_dst:
                    relu_done.add(dst)
                outs[dst].append(out)
                continue
            conv = self.convs[key]
            if streams is not None:          # inputs produced on another stream are read on this relation's stream
                _uses(x_dict[src], rel_streams[edge_type])
                _uses(x_dict[dst], rel_streams[edge_type])
            with torch.cuda.stream(rel_streams[edge_type]) if streams is not None else contextlib.nullcontext():
                if src == dst:
                    out = conv(x_dict[src], edge_index)
                else:
                    out = conv((x_dict[src], x_dict[dst]), edge_index)
            if streams is not None:
                _uses(out, streams[dst])     # summed on the destination type's stream
            outs[dst].append(out)
        result: Dict[str, Tensor] = {}
        for key, xs in outs.items():
            with torch.cuda.stream(streams[key]) if streams is not None else contextlib.nullcontext():
                for r in joins[key]:
                    streams[key].wait_stream(r)
                result[key] = self._aggregate(xs)
                if key in self.fuse_relu_dst and key not in relu_done:
                    result[key] = result[key].relu()
        self.last_streams = streams
        if streams is not None and not self.defer_join:
            for dst in dsts[1:]:
                main.wait_stream(streams[dst])
        return result

    def _aggregate(self, xs: List[Tensor]) -> Tensor:
        if self.aggr in ("sum", "add"):
            acc = xs[0]
            for t in xs[1:]:      # == torch.stack(xs).sum(0) for <= 2 terms; left-to-right beyond
                acc = acc + t
            return acc
        if self.aggr == "mean":
            return torch.stack(xs, 0).mean(0)
        if self.aggr == "max":
            return torch.stack(xs, 0).max(0)[0]
        if self.aggr == "min":
            return torch.stack(xs, 0).min(0)[0]
        if self.aggr == "cat":
            return torch.cat(xs, -1)
        return torch.stack(xs, 1)


class Sequential(nn.Module):
    """torch_geometric.nn.Sequential(input_args, [(module, 'a, b -> c') | callable, ...]) (hscn.py:30-45)."""

    def __init__(self, input_args: str, modules: List[Union[Tuple[Callable, str], Callable]]):
        super().__init__()
        self._inputs = [a.strip() for a in input_args.split(",")]
        self._steps: List[Tuple[str, List[str], List[str]]] = []
        last = [self._inputs[0]]
        for i, entry in enumerate(modules):
            if isinstance(entry, (tuple, list)):
                fn, desc = entry
                lhs, rhs = desc.split("->")
                ins = [a.strip() for a in lhs.split(",")]
                outs = [a.strip() for a in rhs.split(",")]
            else:
                fn, ins, outs = entry, list(last), list(last)
            name = f"module_{i}"
            if isinstance(fn, nn.Module):
                self.add_module(name, fn)
            else:
                object.__setattr__(self, name, fn)
            self._steps.append((name, ins, outs))
            last = outs

    def forward(self, *args):
        env = dict(zip(self._inputs, args))
        out = None
        for name, ins, outs in self._steps:
            out = getattr(self, name)(*[env[k] for k in ins])
            if len(outs) == 1:
                env[outs[0]] = out
            else:
                env.update(zip(outs, out))
        return out
