"""Laplacian positional-encoding precompute on the GPU, batched over graphs.

Mirror of graph_hscn/transform/posenc.py: `compute_posenc_stats(data, is_undirected, cfg)` (:14-47) with the same
argument meaning and the same two attributes set on the object it returns,

    data.eigvals_sn  [N, max_freqs, 1]   the max_freqs smallest Laplacian eigenvalues of the node's graph (clamped at 0)
    data.eigvecs_sn  [N, max_freqs]      the node's entries of the matching eigenvectors, normalised (posenc.py:85-108)

-- but `data` may be a collated Batch: the reference precomputes graph by graph on the host
(loader/loader.py:76-85: get_laplacian -> scipy -> np.linalg.eigh), here every graph of the batch is one CTA of
`ghscn_laplacian_eig` (csrc/posenc.cu).  A single `Data` is a batch of one.  The results land on the device of
`data.edge_index` staged to CUDA; `Batch.to_data_list`-style splitting is a row slice by `ptr`.

Eigenvector signs and the basis inside a repeated eigenvalue are LAPACK's arbitrary choice in the reference; the
consumer (encoder/signnet.py) is sign invariant.  tests/test_gpu_posenc.py checks eigenvalues (2e-5 absolute, the
reference is single precision), residuals ||L v - lambda v||, orthonormality and, for isolated eigenvalues,
|<v, v_ref>| = 1 against oracle/posenc.py.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ._lib import lib
from .structure import _p, _stream, structure_cache

LAPLACIAN_NORMS = {"none": 0, "sym": 1, "rw": 2}
EIGVEC_NORMS = {"L1": 0, "L2": 1, "abs-max": 2}
MAX_SWEEPS = 40


def laplacian_eig(edge_index: Tensor, ptr: Tensor, num_nodes: int, max_nodes_per_graph: int, *,
                  is_undirected: bool = True, laplacian_norm: str = "sym", max_freqs: int = 10,
                  eigvec_norm: str = "L2", return_sweeps: bool = False):
    """-> (eigvals [N, max_freqs], eigvecs [N, max_freqs]) for the graphs delimited by `ptr` (int [B+1])."""
    if not edge_index.is_cuda:
        raise RuntimeError("laplacian_eig is CUDA-only (sm_100a); move the batch to the GPU first")
    norm = LAPLACIAN_NORMS[(laplacian_norm or "none").lower()]
    if eigvec_norm not in EIGVEC_NORMS:
        raise ValueError(f"Unsupported normalization `{eigvec_norm}`")          # posenc.py:104
    dev = edge_index.device
    B = ptr.numel() - 1
    ptr32 = ptr.to(device=dev, dtype=torch.int32).contiguous()
    csr = structure_cache().graph(edge_index, num_nodes, num_nodes, False).by_src   # rows = edge_index[0]
    vals = torch.empty((num_nodes, max_freqs), dtype=torch.float32, device=dev)
    vecs = torch.empty((num_nodes, max_freqs), dtype=torch.float32, device=dev)
    sweeps = torch.empty(B, dtype=torch.int32, device=dev)
    L = lib()
    ws_bytes = L.query("ghscn_laplacian_eig_workspace_bytes", num_nodes, B, int(max_nodes_per_graph))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("ghscn_laplacian_eig", _p(ptr32), _p(csr.rowptr), _p(csr.col), B, num_nodes, int(max_nodes_per_graph),
           norm, int(not is_undirected), int(max_freqs), EIGVEC_NORMS[eigvec_norm], _p(vals), _p(vecs), _p(sweeps),
           _p(ws), ws_bytes, _stream())
    return (vals, vecs, sweeps) if return_sweeps else (vals, vecs)


def _graph_layout(data) -> Tuple[Tensor, int, int]:
    n = int(data.num_nodes) if getattr(data, "num_nodes", None) is not None else int(data.x.shape[0])   # posenc.py:17-20
    ptr = getattr(data, "ptr", None)
    if ptr is None:
        return torch.tensor([0, n], dtype=torch.int64), n, n
    counts = ptr[1:] - ptr[:-1]
    return ptr, n, int(counts.max()) if counts.numel() else 0


def compute_posenc_stats(data, is_undirected: bool, cfg, device: Optional[torch.device] = None):
    """posenc.py:14-47 for a `Data` or a collated `Batch`; `cfg` carries eigen_laplacian_norm, eigen_max_freqs and
    eigvec_norm (config/config.py:124-126)."""
    ptr, n, max_nodes = _graph_layout(data)
    edge_index = data.edge_index
    if not edge_index.is_cuda:
        edge_index = edge_index.to(device or torch.device("cuda", torch.cuda.current_device()))
    vals, vecs, sweeps = laplacian_eig(edge_index, ptr, n, max_nodes, is_undirected=is_undirected,
                                       laplacian_norm=cfg.eigen_laplacian_norm, max_freqs=cfg.eigen_max_freqs,
                                       eigvec_norm=cfg.eigvec_norm, return_sweeps=True)
    # a one-time precompute: one host read to fail loudly instead of storing eigenpairs of an iteration that did not
    # converge (never observed: 6-15 sweeps on Peptides / VOC shapes against a cap of 40)
    if sweeps.numel() and int(sweeps.max()) >= MAX_SWEEPS:
        raise RuntimeError("ghscn_laplacian_eig: the Jacobi iteration did not converge for at least one graph")
    data.eigvals_sn, data.eigvecs_sn = vals.unsqueeze(2), vecs
    return data
