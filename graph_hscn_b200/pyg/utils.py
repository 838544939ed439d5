"""`torch_geometric.utils` names the reference's PE transform imports (transform/posenc.py:5-9): host-side index glue
with PyG 2.2/2.3 semantics, device agnostic, so `graph_hscn/transform/posenc.py` imports and runs unchanged under
`pyg.install()` (its `np.linalg.eigh` host loop included).  The accelerated replacement of that whole function is
`graph_hscn_b200.posenc.compute_posenc_stats` (one CTA per graph of `ghscn_laplacian_eig`); these helpers are not on its
path -- the kernel builds the dense Laplacian from the CSR slice itself.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tuple[Tensor, Optional[Tensor]]:
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (edge_attr[keep] if edge_attr is not None else None)


def to_undirected(edge_index: Tensor, edge_attr: Optional[Tensor] = None, num_nodes: Optional[int] = None) -> Tensor:
    """Both directions of every edge, coalesced (sorted by (row, col), duplicates dropped)."""
    if edge_attr is not None and not isinstance(edge_attr, int):
        raise NotImplementedError("to_undirected with edge attributes is not on the reference's path")
    if isinstance(edge_attr, int):
        num_nodes = edge_attr
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    both = torch.cat([edge_index, edge_index.flip(0)], 1)
    key = torch.unique(both[0] * n + both[1])
    return torch.stack([torch.div(key, n, rounding_mode="floor"), key % n])


def get_laplacian(edge_index: Tensor, edge_weight: Optional[Tensor] = None, normalization: Optional[str] = None,
                  dtype: Optional[torch.dtype] = None, num_nodes: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """(edge_index, weights) of L = D - A | I - D^-1/2 A D^-1/2 ('sym') | I - D^-1 A ('rw'): the off-diagonal entries
    first, then one diagonal entry per node."""
    if normalization not in (None, "sym", "rw"):
        raise ValueError(f"Invalid normalization {normalization!r}")
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    dev = edge_index.device
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=dtype or torch.float32, device=dev)
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    row, col = edge_index
    deg = torch.zeros(n, dtype=edge_weight.dtype, device=dev).scatter_add_(0, row, edge_weight)
    diag = torch.arange(n, dtype=edge_index.dtype, device=dev)
    index = torch.cat([edge_index, torch.stack([diag, diag])], 1)
    if normalization is None:
        return index, torch.cat([-edge_weight, deg])
    if normalization == "sym":
        scale = deg.pow(-0.5)
        scale.masked_fill_(scale == float("inf"), 0)
        off = scale[row] * edge_weight * scale[col]
    else:
        scale = 1.0 / deg
        scale.masked_fill_(scale == float("inf"), 0)
        off = scale[row] * edge_weight
    return index, torch.cat([-off, torch.ones(n, dtype=edge_weight.dtype, device=dev)])


def to_scipy_sparse_matrix(edge_index: Tensor, edge_attr: Optional[Tensor] = None, num_nodes: Optional[int] = None):
    import scipy.sparse
    row, col = edge_index.detach().cpu().numpy()
    vals = torch.ones(row.shape[0]) if edge_attr is None else edge_attr.detach().view(-1).cpu()
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    return scipy.sparse.coo_matrix((vals.numpy(), (row, col)), (n, n))
