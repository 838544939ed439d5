"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Laplacian positional-encoding precompute.

Follows, line by line,
    graph_hscn/transform/posenc.py:14-47   compute_posenc_stats  (Laplacian -> dense -> np.linalg.eigh)
    graph_hscn/transform/posenc.py:50-82   get_lap_decomp_stats  (k smallest, normalise, NaN padding)
    graph_hscn/transform/posenc.py:85-108  eigvec_normalizer
and PyG 2.2/2.3's `get_laplacian`, `to_undirected`, `remove_self_loops`, `to_scipy_sparse_matrix` (not vendored in
/root/reference; restated from their published behaviour -- "parity unpinned" for those four, like the rest of the PyG
operator layer).  `get_lap_decomp_stats` / `eigvec_normalizer` ARE pinned: tests/golden/posenc.pt holds the outputs of
the reference's own source text for them (tests/golden/make_golden_posenc.py executes it unmodified), and the rows that
the reference's whole `compute_posenc_stats`, imported unchanged, produces on top of the four restated utilities.

Precision note: PyG builds the Laplacian weights in float32 and scipy's `.toarray()` keeps that dtype, so the
reference's `np.linalg.eigh` is LAPACK single precision (ssyevd).  Eigenvector signs, and the basis inside a repeated
eigenvalue, are whatever LAPACK returns: parity for them is defined up to sign / up to the invariant subspace.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


def remove_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor] = None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_weight is None else edge_weight[mask])


def to_undirected(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """PyG to_undirected without edge_attr: cat with the flipped edges, then coalesce (sort by row*N+col, drop
    duplicates)."""
    n = int(edge_index.max()) + 1 if num_nodes is None and edge_index.numel() else int(num_nodes or 0)
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    key = torch.unique(row * n + col)                 # sorted, deduplicated
    return torch.stack([key // n, key % n])


def get_laplacian(edge_index: Tensor, edge_weight: Optional[Tensor] = None, normalization: Optional[str] = None,
                  num_nodes: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """PyG get_laplacian: L = D - A (None), I - D^-1/2 A D^-1/2 ('sym'), I - D^-1 A ('rw'), as COO with the
    diagonal appended behind the off-diagonal entries; float32 weights."""
    assert normalization in (None, "sym", "rw")
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1), dtype=torch.float32)
    n = int(num_nodes)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(n, dtype=edge_weight.dtype).scatter_add_(0, row, edge_weight)
    loops = torch.arange(n, dtype=edge_index.dtype)
    loop_index = torch.stack([loops, loops])
    if normalization is None:
        edge_index = torch.cat([edge_index, loop_index], 1)
        edge_weight = torch.cat([-edge_weight, deg])
    elif normalization == "sym":
        dis = deg.pow(-0.5)
        dis.masked_fill_(dis == float("inf"), 0)
        edge_weight = dis[row] * edge_weight * dis[col]
        edge_index = torch.cat([edge_index, loop_index], 1)
        edge_weight = torch.cat([-edge_weight, torch.ones(n, dtype=edge_weight.dtype)])
    else:
        dinv = 1.0 / deg
        dinv.masked_fill_(dinv == float("inf"), 0)
        edge_weight = dinv[row] * edge_weight
        edge_index = torch.cat([edge_index, loop_index], 1)
        edge_weight = torch.cat([-edge_weight, torch.ones(n, dtype=edge_weight.dtype)])
    return edge_index, edge_weight


def to_scipy_sparse_matrix(edge_index: Tensor, edge_attr: Optional[Tensor] = None, num_nodes: Optional[int] = None):
    """PyG to_scipy_sparse_matrix: COO matrix of the edge list (duplicates add up on `.toarray()`), dtype of the
    weights (float32 for a Laplacian from get_laplacian)."""
    import scipy.sparse
    row, col = edge_index.cpu().numpy()
    if edge_attr is None:
        edge_attr = torch.ones(row.shape[0])
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    return scipy.sparse.coo_matrix((edge_attr.view(-1).cpu().numpy(), (row, col)), (n, n))


def laplacian_dense(edge_index: Tensor, num_nodes: int, is_undirected: bool, norm: Optional[str]) -> np.ndarray:
    """posenc.py:30-41: `to_scipy_sparse_matrix(*get_laplacian(...)).toarray()` -- duplicate entries add up, dtype
    float32."""
    norm = None if norm is None or norm.lower() == "none" else norm.lower()
    und = edge_index if is_undirected else to_undirected(edge_index, num_nodes)
    ei, ew = get_laplacian(und, normalization=norm, num_nodes=num_nodes)
    dense = torch.zeros(num_nodes * num_nodes, dtype=torch.float32)
    dense.scatter_add_(0, ei[0] * num_nodes + ei[1], ew)
    return dense.view(num_nodes, num_nodes).numpy()


def eigvec_normalizer(eig_vecs: Tensor, eig_vals: Tensor, normalization: str = "L2", eps: float = 1e-12) -> Tensor:
    """posenc.py:85-108."""
    if normalization == "L1":
        denom = eig_vecs.norm(p=1, dim=0, keepdim=True)
    elif normalization == "L2":
        denom = eig_vecs.norm(p=2, dim=0, keepdim=True)
    elif normalization == "abs-max":
        denom = torch.max(eig_vecs.abs(), dim=0, keepdim=True).values
    else:
        raise ValueError(f"Unsupported normalization `{normalization}`")
    denom = denom.clamp_min(eps).expand_as(eig_vecs)
    return eig_vecs / denom


def get_lap_decomp_stats(evals: np.ndarray, evects: np.ndarray, max_freqs: int,
                         eigvec_norm: str = "L2") -> Tuple[Tensor, Tensor]:
    """posenc.py:50-82 -> (eigvals [N, max_freqs, 1], eigvecs [N, max_freqs]), NaN-padded when N < max_freqs."""
    n = len(evals)
    idx = evals.argsort()[:max_freqs]
    evals, evects = evals[idx], np.real(evects[:, idx])
    evals_t = torch.from_numpy(np.real(evals)).clamp_min(0)
    evects_t = torch.from_numpy(evects).float()
    evects_t = eigvec_normalizer(evects_t, evals_t, normalization=eigvec_norm)
    if n < max_freqs:
        eig_vecs = F.pad(evects_t, (0, max_freqs - n), value=float("nan"))
        eig_vals = F.pad(evals_t, (0, max_freqs - n), value=float("nan")).unsqueeze(0)
    else:
        eig_vecs = evects_t
        eig_vals = evals_t.unsqueeze(0)
    eig_vals = eig_vals.repeat(n, 1).unsqueeze(2)
    return eig_vals, eig_vecs


def compute_posenc_stats(edge_index: Tensor, num_nodes: int, is_undirected: bool, max_freqs: int = 10,
                         eigvec_norm: str = "L2", laplacian_norm: str = "sym") -> Tuple[Tensor, Tensor]:
    """posenc.py:14-47 for one graph -> (eigvals_sn, eigvecs_sn)."""
    lap = laplacian_dense(edge_index, num_nodes, is_undirected, laplacian_norm)
    evals, evects = np.linalg.eigh(lap)
    return get_lap_decomp_stats(evals, evects, max_freqs, eigvec_norm)
