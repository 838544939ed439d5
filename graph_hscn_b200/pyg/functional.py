"""PyG-named functional operators backed by the sm_100a kernels (drop-in boundary).

Signatures, argument meaning and error behaviour follow torch_geometric / torch_scatter for the call
sites of the reference:
    gcn_norm            train/train_clustering.py:37-42,58-63
    to_dense_adj        model/hscn.py:61
    dense_mincut_pool   model/hscn.py:63
    global_mean_pool    model/hscn.py:111
    scatter_mean        model/mpnn.py:60
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import ops
from .._lib import lib
from . import _device
from ..structure import (GraphStructure, Segments, _capturing, _p, _require_cuda, _stream, current_hints,
                         structure_cache)


# ---------------------------------------------------------------------------------------------
# gcn_norm  
# ---------------------------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor] = None, num_nodes: Optional[int] = None,
             improved: bool = False, add_self_loops: bool = True, flow: str = "source_to_target",
             dtype: Optional[torch.dtype] = None) -> Tuple[Tensor, Tensor]:
    """Returns PyG's (edge_index', edge_weight') in PyG's edge order (original non-loop edges, then the
    N appended loops).  The layers never call this -- they consume the CSR directly -- it exists for the
    reference's direct call (train_clustering.py:37) and is computed by the same kernels.  The reference calls it
    on CPU tensors (before `data.to(device)`): with auto-device staging they are copied to the GPU, the kernels run
    there and both results come back on the input's device (pyg/_device.py)."""
    dev, back = _device.plan(edge_index, edge_weight)
    if dev is not None:
        edge_index, edge_weight = _device.index_to_dev(edge_index, dev), _device.to_dev(edge_weight, dev)
    return _device.back_to(_gcn_norm_cuda(edge_index, edge_weight, num_nodes, improved, add_self_loops, flow), back)


def _gcn_norm_cuda(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: Optional[int], improved: bool,
                   add_self_loops: bool, flow: str) -> Tuple[Tensor, Tensor]:
    _require_cuda(edge_index, edge_weight)
    if flow != "source_to_target":
        raise NotImplementedError("only flow='source_to_target' is on the reference's path")
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() else 0
    st = structure_cache().graph(edge_index, num_nodes, num_nodes, add_self_loops)
    w, _, _ = st.weights(edge_weight, normalize=True, improved=improved, need_transpose=False)
    d = st.by_dst
    M = d.num_items
    no_drop = (not add_self_loops) or bool(current_hints().get("no_self_loops"))
    if not no_drop:
        if _capturing():
            raise RuntimeError("gcn_norm must know whether self loops exist; use structure_hints(no_self_loops=1)")
        no_drop = int(d.rowptr[-1].item()) == M   # PyG's masked indexing syncs here as well
    if no_drop:
        # nothing was dropped: item ids are exactly the positions of PyG's output edge list
        if add_self_loops:
            loops = torch.arange(num_nodes, device=edge_index.device).unsqueeze(0).expand(2, num_nodes)
            new_index = torch.cat([edge_index, loops], dim=1)
        else:
            new_index = edge_index
        w_coo = torch.empty(M, dtype=torch.float32, device=edge_index.device)
        w_coo[d.perm.long()] = w
        if add_self_loops:   # layers called with (new_index, w_coo) reuse this CSR instead of sorting again ...
            view = structure_cache().alias(new_index, st)
            # ... and its per-slot weights: w_coo gathered back to slot order is `w` itself
            view._weights[((w_coo.data_ptr(), w_coo._version), False, False)] = dict(
                w=w, w_t=None, dis=None, lw=None, ew=w_coo, unit=False)
        return new_index, w_coo
    nnz = int(d.rowptr[-1].item())
    # back to PyG's COO order: slot s holds item perm[s]; items = kept edges (in order) then loops
    perm = d.perm[:nnz].long()
    rows = d.col[:nnz].long()           # `col` of the by-destination CSR is the source endpoint
    dst_of_slot = torch.repeat_interleave(torch.arange(num_nodes, device=edge_index.device),
                                          (d.rowptr[1:] - d.rowptr[:-1]).long())
    order = torch.argsort(perm, stable=True)
    new_index = torch.stack([rows[order], dst_of_slot[order]])
    return new_index, w[:nnz][order]


# ---------------------------------------------------------------------------------------------
# to_dense_adj -> lazy CSR-backed adjacency  ("dense API / sparse implementation")
# ---------------------------------------------------------------------------------------------
class SparseAdj:
    """What `to_dense_adj` returns: the batch adjacency kept as CSR.  `dense_mincut_pool` consumes it
    directly; any other use densifies it (`.to_dense()`, or implicitly through torch functions)."""

    def __init__(self, edge_index: Tensor, batch: Optional[Tensor], edge_attr: Optional[Tensor],
                 max_num_nodes: Optional[int]):
        dev, self._back = _device.plan(edge_index, batch, edge_attr)
        if dev is not None:
            edge_index, batch = _device.index_to_dev(edge_index, dev), _device.index_to_dev(batch, dev)
            edge_attr = _device.to_dev(edge_attr, dev)
        self.edge_index, self.batch, self.edge_attr, self.max_num_nodes = edge_index, batch, edge_attr, max_num_nodes
        self._dense: Optional[Tensor] = None

    @property
    def device(self):
        return self.edge_index.device

    def to_dense(self) -> Tensor:
        if self._dense is None:
            ei, batch = self.edge_index, self.batch
            if batch is None:
                n = int(ei.max()) + 1 if ei.numel() else 0
                batch = ei.new_zeros(n)
            B = int(batch.max()) + 1 if batch.numel() else 1
            counts = torch.bincount(batch, minlength=B)
            cum = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
            n_max = int(counts.max()) if self.max_num_nodes is None else self.max_num_nodes
            i0 = batch[ei[0]]
            i1 = ei[0] - cum[i0]
            i2 = ei[1] - cum[batch[ei[1]]]
            val = self.edge_attr if self.edge_attr is not None else torch.ones(ei.size(1), device=ei.device)
            keep = (i1 < n_max) & (i2 < n_max)
            flat = torch.zeros(B * n_max * n_max, dtype=val.dtype, device=ei.device)
            flat.index_add_(0, (i0 * n_max * n_max + i1 * n_max + i2)[keep], val[keep])
            self._dense = _device.back_to(flat.view(B, n_max, n_max), self._back)
        return self._dense

    def size(self, dim: Optional[int] = None):
        return self.to_dense().size() if dim is None else self.to_dense().size(dim)

    @property
    def shape(self):
        return self.to_dense().shape

    def dim(self) -> int:
        return 3

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        conv = lambda a: a.to_dense() if isinstance(a, SparseAdj) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})


def to_dense_adj(edge_index: Tensor, batch: Optional[Tensor] = None, edge_attr: Optional[Tensor] = None,
                 max_num_nodes: Optional[int] = None) -> SparseAdj:
    return SparseAdj(edge_index, batch, edge_attr, max_num_nodes)


# ---------------------------------------------------------------------------------------------
# MinCUT pool  
# ---------------------------------------------------------------------------------------------
def _ptr_for(batch: Optional[Tensor], num_nodes: int, device) -> Tuple[Tensor, int, int]:
    """-> (ptr int32 [B+1], B, max nodes per graph)"""
    if batch is None:
        ptr = torch.tensor([0, num_nodes], dtype=torch.int32, device=device)
        return ptr, 1, max(num_nodes, 1)
    seg = structure_cache().segments(batch)
    max_rows = seg.max_rows or current_hints().get("max_nodes_per_graph", 0)
    if not max_rows:
        if _capturing():
            raise RuntimeError("max nodes per graph needs a host sync; use structure_hints(max_nodes_per_graph=)")
        max_rows = int((seg.ptr[1:] - seg.ptr[:-1]).max().item()) if seg.num_segments else 1
        seg.max_rows = max_rows
    return seg.ptr, seg.num_segments, max(max_rows, 1)


def mincut_pool_ragged(x: Tensor, edge_index: Tensor, s: Tensor, batch: Optional[Tensor] = None,
                       edge_attr: Optional[Tensor] = None, temp: float = 1.0, want_out: bool = True,
                       want_adj: bool = True, losses_tensor: bool = False, num_graphs: Optional[int] = None):
    """Batched MinCUT pool on the ragged (`batch`/`ptr`) layout: one CTA per graph, no [B,n,n] tensor.
    Equivalent to to_dense_batch + to_dense_adj(batch) + dense_mincut_pool(mask) in PyG.
    -> (out, out_adj, mincut_loss, ortho_loss), or with `losses_tensor` (out, out_adj, losses[2]) so that a caller
    that only adds the two losses does not pay for two select/scatter round trips in autograd.
    `num_graphs` pools only the first that many graphs of the batch (the rest -- trailing padding graphs of a
    bucketed batch -- get no output rows, no share of the loss means and zero gradients)."""
    _require_cuda(x, edge_index, s)
    N = s.size(0)
    ptr, B, max_nodes = _ptr_for(batch, N, s.device)
    if num_graphs is not None and num_graphs < B:
        ptr, B = ptr[:num_graphs + 1], int(num_graphs)
    st = structure_cache().graph(edge_index, N, N, False)
    rows, cols = st.by_src, st.by_dst   # rows of A are edge_index[0]
    val = val_t = None
    if edge_attr is not None:
        val, val_t, _ = st.weights(edge_attr, normalize=False)
        val, val_t = val_t, val         # by_src carries A's rows
    out, out_adj, losses, *_ = torch.ops.ghscn.mincut_pool(
        s, x, ptr, rows.rowptr, rows.col, val, cols.rowptr, cols.col, val_t, float(temp), int(max_nodes),
        bool(want_out), bool(want_adj))
    if losses_tensor:
        return (out if want_out else None), (out_adj if want_adj else None), losses
    return (out if want_out else None), (out_adj if want_adj else None), losses[0], losses[1]


def dense_mincut_pool(x: Tensor, adj, s: Tensor, mask: Optional[Tensor] = None, temp: float = 1.0
                      ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """PyG signature and returns: (out [B,K,H], out_adj [B,K,K], mincut_loss, ortho_loss)."""
    dev, back = _device.plan(x, s, adj if isinstance(adj, Tensor) else adj.edge_index, mask)
    if dev is not None:
        x, s, mask = _device.to_dev(x, dev), _device.to_dev(s, dev), _device.to_dev(mask, dev)
        adj = _device.to_dev(adj, dev)          # a SparseAdj staged its own tensors when it was built
    return _device.back_to(_dense_mincut_pool_cuda(x, adj, s, mask, temp), back)


def _dense_mincut_pool_cuda(x: Tensor, adj, s: Tensor, mask: Optional[Tensor], temp: float):
    _require_cuda(x, s)
    if isinstance(adj, SparseAdj) and x.dim() == 2 and mask is None:
        return mincut_pool_ragged(x, adj.edge_index, s, adj.batch, adj.edge_attr, temp)
    # dense-input overload: densify-free conversion of the given adjacency to COO, then the same kernel
    dense = adj.to_dense() if isinstance(adj, SparseAdj) else adj
    dense = dense.unsqueeze(0) if dense.dim() == 2 else dense
    x3 = x.unsqueeze(0) if x.dim() == 2 else x
    s3 = s.unsqueeze(0) if s.dim() == 2 else s
    B, n, _ = x3.shape
    if mask is None:
        keep = torch.ones(B, n, dtype=torch.bool, device=x.device)
    else:
        keep = mask.view(B, n).bool()
    counts = keep.sum(1)
    new_id = (torch.cumsum(keep.view(-1).long(), 0) - 1).view(B, n)     # ragged node ids
    b_idx, r_idx, c_idx = dense.nonzero(as_tuple=True)
    ok = keep[b_idx, r_idx] & keep[b_idx, c_idx]
    if not bool(ok.all()):
        # PyG would count edges into masked nodes in the degree term; padded nodes of a collated batch never
        # have edges, so refuse instead of silently computing something else.
        raise ValueError("dense_mincut_pool: adjacency has entries on masked-out nodes")
    ei = torch.stack([new_id[b_idx, r_idx], new_id[b_idx, c_idx]])
    vals = dense[b_idx, r_idx, c_idx].float()
    batch = torch.repeat_interleave(torch.arange(B, device=x.device), counts)
    xr = x3[keep]
    sr = s3[keep]
    with_hint = {"num_graphs": B, "batch_sorted": 1, "max_nodes_per_graph": n}
    from ..structure import structure_hints
    with structure_hints(**with_hint):
        return mincut_pool_ragged(xr, ei, sr, batch, vals, temp)


# ---------------------------------------------------------------------------------------------
# readout  
# ---------------------------------------------------------------------------------------------
def _segment(x: Tensor, index: Tensor, size: Optional[int], mean: bool) -> Tensor:
    dev, back = _device.plan(x, index)
    if dev is not None:
        return _device.back_to(_segment(_device.to_dev(x, dev), _device.index_to_dev(index, dev), size, mean), back)
    _require_cuda(x, index)
    squeeze = x.dim() == 1
    x2 = x.unsqueeze(1) if squeeze else x
    if x2.dtype != torch.float32:
        x2 = x2.float()
    seg = structure_cache().segments(index, size)
    out = torch.ops.ghscn.segment_reduce(x2, seg.ptr, seg.perm, mean)
    return out.squeeze(1) if squeeze else out


def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.mean(dim=-2, keepdim=x.dim() == 2)
    return _segment(x, batch, size, True)


def global_add_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.sum(dim=-2, keepdim=x.dim() == 2)
    return _segment(x, batch, size, False)


def scatter_mean(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                 dim_size: Optional[int] = None) -> Tensor:
    """torch_scatter.scatter_mean for the layout the reference uses: 1-D index over dim 0 of [N,F]."""
    if out is not None:
        raise NotImplementedError("scatter_mean(out=...) is not on the reference's path")
    if dim < 0:
        dim += src.dim()
    if dim != 0 or index.dim() != 1 or src.dim() > 2:
        raise NotImplementedError("only scatter_mean(x [N,F], batch [N], dim=0) is on the reference's path")
    return _segment(src, index, dim_size, True)


def scatter_sum(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    if out is not None:
        raise NotImplementedError
    if dim < 0:
        dim += src.dim()
    if dim != 0 or index.dim() != 1 or src.dim() > 2:
        raise NotImplementedError("only scatter(x [N,F], index [N], dim=0) is supported")
    return _segment(src, index, dim_size, False)


scatter_add = scatter_sum


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    raise NotImplementedError(f"scatter reduce={reduce!r} is not on the reference's path")


def scn_logits_fused(x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor], conv, act: str, out_lin
                     ) -> Optional[Tensor]:
    """Cluster logits of SCN(mp_units=[U]) -- GraphConv `conv`, activation `act`, Linear `out_lin`
    (model/hscn.py:30-45,57-60) -- in one launch (ghscn_scn_forward); None when the shape or dtype is outside the
    fused kernel's range (the caller then runs the separate operators)."""
    from .. import ops
    from ..structure import structure_cache
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and act in ops.SCN_ACTS):
        return None
    if x.requires_grad or (edge_weight is not None and edge_weight.requires_grad):
        return None             # the fused op differentiates w.r.t. the parameters only
    f, u, k = x.size(1), conv.out_channels, out_lin.out_channels
    if f > ops.SCN_LIMITS[0] or u > ops.SCN_LIMITS[1] or k > ops.SCN_LIMITS[2]:
        return None
    if conv.lin_rel.weight.size(1) != f or conv.lin_root.weight.size(1) != f or out_lin.weight.size(1) != u:
        return None
    st = structure_cache().graph(edge_index, x.size(0), x.size(0), False)
    w, _, _ = st.weights(edge_weight, normalize=False, need_transpose=False)
    d = st.by_dst
    xs = x if x.stride(1) == 1 else x.contiguous()
    return ops.scn_node_forward(xs, d.rowptr, d.col, w, conv.lin_rel.weight, conv.lin_rel.bias, conv.lin_root.weight,
                                out_lin.weight, out_lin.bias, act)


def linear_act(lin, x: Tensor, act) -> Optional[Tensor]:
    """act(lin(x)) for a `Linear` module in one launch (bias and activation in the GEMM epilogue) when the operator
    set can fuse it; None otherwise.  `act` is a callable from config/config.py:13-18 (F.relu, F.elu, torch.tanh)."""
    import torch.nn.functional as F
    from .. import gemm
    name = {F.relu: "relu", torch.relu: "relu", F.elu: "elu", torch.tanh: "tanh", F.tanh: "tanh"}.get(act)
    if isinstance(act, torch.nn.Identity):
        name = "identity"
    if name is None or not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32:
        return None
    lin.materialize(x.size(-1), x.device)
    if not lin.weight.is_cuda:
        return None
    return gemm.linear_act(x, lin.weight, lin.bias, name)
