"""Functional restatement of the PyG / torch_scatter operators on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference call site it serves and the SURVEY.md appendix paragraph that states
the third-party semantics being restated.  `row = edge_index[0]` is the source
j, `col = edge_index[1]` the target i (flow source_to_target); all edge sums
run in edge order through `scatter_add_` exactly like PyG on CPU.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor


# --------------------------------------------------------------------------
# torch_scatter restatement (mpnn.py:8,60 ; used inside PyG <= 2.2)
# --------------------------------------------------------------------------
def _broadcast(index: Tensor, src: Tensor, dim: int) -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                 dim_size: Optional[int] = None) -> Tensor:
    """torch_scatter.scatter_mean as called at mpnn.py:60 (`scatter_mean(x, batch, dim=0)`)."""
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter_max(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None) -> Tensor:
    index_b = _broadcast(index, src, dim)
    size = list(src.size())
    size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    out = torch.full(size, float("-inf"), dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, index_b, src, reduce="amax", include_self=True)
    # torch_scatter leaves empty segments at 0
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)


def scatter(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, dim_size)
    raise ValueError(reduce)


# --------------------------------------------------------------------------
# A.1 / A.2  gcn_norm  (train_clustering.py:37-42,58-63 ; inside every GCNConv)
# --------------------------------------------------------------------------
def maybe_num_nodes(edge_index: Tensor, num_nodes: Optional[int] = None) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def add_remaining_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                             fill_value: float = 1.0, num_nodes: Optional[int] = None
                             ) -> Tuple[Tensor, Optional[Tensor]]:
    """drop existing loops, append N loops; an existing loop keeps its weight."""
    N = maybe_num_nodes(edge_index, num_nodes)
    mask = edge_index[0] != edge_index[1]
    loop_index = torch.arange(0, N, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        loop_attr = edge_attr.new_full((N,) + tuple(edge_attr.size()[1:]), fill_value)
        inv_mask = ~mask
        loop_attr[edge_index[0][inv_mask]] = edge_attr[inv_mask]
        edge_attr = torch.cat([edge_attr[mask], loop_attr], dim=0)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, edge_attr


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor] = None, num_nodes: Optional[int] = None,
             improved: bool = False, add_self_loops: bool = True, flow: str = "source_to_target",
             dtype: Optional[torch.dtype] = None) -> Tuple[Tensor, Tensor]:
    """SURVEY A.2.  Multiplication order dis[row] * w * dis[col] is part of the contract."""
    fill_value = 2.0 if improved else 1.0
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, num_nodes)
    row, col = edge_index[0], edge_index[1]
    idx = col if flow == "source_to_target" else row
    deg = scatter_sum(edge_weight, idx, dim=0, dim_size=num_nodes)
    deg_inv_sqrt = deg.pow_(-0.5)
    deg_inv_sqrt.masked_fill_(deg_inv_sqrt == float("inf"), 0)
    return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]


def propagate_add(x_src: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor], num_dst: int) -> Tensor:
    """MessagePassing.propagate with aggr='add': gather -> (scale) -> scatter_add_ in edge order."""
    row, col = edge_index[0], edge_index[1]
    msg = x_src.index_select(0, row)
    if edge_weight is not None:
        msg = edge_weight.view(-1, 1) * msg
    return scatter_sum(msg, col, dim=0, dim_size=num_dst)


# --------------------------------------------------------------------------
# A.5  to_dense_adj  (hscn.py:61)
# --------------------------------------------------------------------------
def to_dense_adj(edge_index: Tensor, batch: Optional[Tensor] = None, edge_attr: Optional[Tensor] = None,
                 max_num_nodes: Optional[int] = None) -> Tensor:
    if batch is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
        batch = edge_index.new_zeros(num_nodes)
    batch_size = int(batch.max()) + 1 if batch.numel() > 0 else 1
    one = batch.new_ones(batch.size(0))
    num_nodes = scatter_sum(one, batch, dim=0, dim_size=batch_size)
    cum_nodes = torch.cat([batch.new_zeros(1), num_nodes.cumsum(dim=0)])
    idx0 = batch[edge_index[0]]
    idx1 = edge_index[0] - cum_nodes[batch][edge_index[0]]
    idx2 = edge_index[1] - cum_nodes[batch][edge_index[1]]
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max()) if num_nodes.numel() else 0
    elif (idx1.numel() > 0 and idx1.max() >= max_num_nodes) or (idx2.numel() > 0 and idx2.max() >= max_num_nodes):
        mask = (idx1 < max_num_nodes) & (idx2 < max_num_nodes)
        idx0, idx1, idx2 = idx0[mask], idx1[mask], idx2[mask]
        edge_attr = None if edge_attr is None else edge_attr[mask]
    if edge_attr is None:
        edge_attr = torch.ones(idx0.numel(), device=edge_index.device)
    size = [batch_size, max_num_nodes, max_num_nodes] + list(edge_attr.size())[1:]
    flattened = batch_size * max_num_nodes * max_num_nodes
    idx = idx0 * max_num_nodes * max_num_nodes + idx1 * max_num_nodes + idx2
    adj = scatter_sum(edge_attr, idx, dim=0, dim_size=flattened)
    return adj.view(size)


def to_dense_batch(x: Tensor, batch: Tensor, max_num_nodes: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    B = int(batch.max()) + 1 if batch.numel() else 1
    counts = torch.bincount(batch, minlength=B)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    n_max = int(counts.max()) if max_num_nodes is None else max_num_nodes
    local = torch.arange(batch.numel()) - ptr[batch]
    out = x.new_zeros((B, n_max) + tuple(x.shape[1:]))
    out[batch, local] = x
    mask = torch.zeros(B, n_max, dtype=torch.bool)
    mask[batch, local] = True
    return out, mask


# --------------------------------------------------------------------------
# A.6  dense_mincut_pool  (hscn.py:63)
# --------------------------------------------------------------------------
def dense_mincut_pool(x: Tensor, adj: Tensor, s: Tensor, mask: Optional[Tensor] = None,
                      temp: float = 1.0) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    x = x.unsqueeze(0) if x.dim() == 2 else x
    adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
    s = s.unsqueeze(0) if s.dim() == 2 else s
    (batch_size, num_nodes, _), k = x.size(), s.size(-1)
    s = torch.softmax(s / temp if temp != 1.0 else s, dim=-1)
    if mask is not None:
        mask = mask.view(batch_size, num_nodes, 1).to(x.dtype)
        x, s = x * mask, s * mask
    out = torch.matmul(s.transpose(1, 2), x)
    out_adj = torch.matmul(torch.matmul(s.transpose(1, 2), adj), s)
    # MinCut regularisation
    mincut_num = torch.einsum("ijj->i", out_adj)
    d_flat = torch.einsum("ijk->ij", adj)
    d = torch.diag_embed(d_flat)
    mincut_den = torch.einsum("ijj->i", torch.matmul(torch.matmul(s.transpose(1, 2), d), s))
    mincut_loss = -(mincut_num / mincut_den)
    mincut_loss = torch.mean(mincut_loss)
    # Orthogonality regularisation
    ss = torch.matmul(s.transpose(1, 2), s)
    i_s = torch.eye(k).type_as(ss)
    ortho_loss = torch.norm(
        ss / torch.norm(ss, dim=(-1, -2), keepdim=True) - i_s / torch.norm(i_s), dim=(-1, -2))
    ortho_loss = torch.mean(ortho_loss)
    EPS = 1e-15
    # Fix and normalise coarsened adjacency
    ind = torch.arange(k, device=out_adj.device)
    out_adj = out_adj.clone()  # autograd-safe equivalent of PyG's in-place diagonal zeroing
    out_adj[:, ind, ind] = 0
    d = torch.einsum("ijk->ij", out_adj)
    d = torch.sqrt(d)[:, None] + EPS
    out_adj = (out_adj / d) / d.transpose(1, 2)
    return out, out_adj, mincut_loss, ortho_loss


# --------------------------------------------------------------------------
# A.9  readout  (hscn.py:111)
# --------------------------------------------------------------------------
def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.mean(dim=-2, keepdim=x.dim() == 2)
    size = int(batch.max().item() + 1) if size is None else size
    return scatter_mean(x, batch, dim=-2, dim_size=size)


def global_add_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.sum(dim=-2, keepdim=x.dim() == 2)
    size = int(batch.max().item() + 1) if size is None else size
    return scatter_sum(x, batch, dim=-2, dim_size=size)


# --------------------------------------------------------------------------
# A.8  segment softmax used by GATConv  (hscn.py:85-87 via CONV_DICT["gat"])
# --------------------------------------------------------------------------
def segment_softmax(src: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    src_max = scatter_max(src.detach(), index, dim=0, dim_size=num_nodes)
    out = src - src_max.index_select(0, index)
    out = out.exp()
    out_sum = scatter_sum(out, index, dim=0, dim_size=num_nodes) + 1e-16
    return out / out_sum.index_select(0, index)


# --------------------------------------------------------------------------
# sparse identities of A.6 (used by tests to cross-check dense vs CSR form)
# --------------------------------------------------------------------------
def mincut_sparse_terms(edge_index: Tensor, s_soft: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """num = sum S*(AS), den = sum_n d_n sum_k S_nk^2, AS -- for one graph, binary multi-edge A."""
    row, col = edge_index[0], edge_index[1]
    n = s_soft.size(0)
    AS = scatter_sum(s_soft.index_select(0, col), row, dim=0, dim_size=n)
    d = scatter_sum(torch.ones(row.numel(), dtype=s_soft.dtype), row, dim=0, dim_size=n)
    num = (s_soft * AS).sum()
    den = (d.view(-1, 1) * s_soft * s_soft).sum()
    return num, den, AS
