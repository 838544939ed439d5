"""Seeded synthetic graph batches shaped like the reference's datasets.

There is no network, ogb or rdkit here, so benchmark and parity inputs are
generated, on the CPU, to match the shapes of
  * Peptides-func / Peptides-struct (loader/dataset/peptides_functional.py:21-115,
    peptides_structural.py:21-121): ~151 nodes, ~307 directed edges per graph,
    9 integer OGB atom features, [1,10] multi-label or [1,11] z-scored targets;
    edges stored as adjacent (i,j),(j,i) pairs in ogb smiles2graph order, unsorted;
  * PascalVOC-SP superpixel graphs: 395-500 nodes, avg degree ~5.66, 14 float features;
  * Erdos-Renyi-per-graph sweeps with a chosen average degree.
The same tensors are fed to the CPU oracle and to the CUDA path.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .data import Batch, Data

ATOM_FEATURE_RANGES = (119, 5, 12, 12, 10, 6, 6, 2, 2)  # OGB atom feature cardinalities


def _peptide_edges(n: int, rng: np.random.Generator) -> np.ndarray:
    """Random degree-capped tree plus a few ring closures; returns [2, E] with (i,j),(j,i) adjacent."""
    deg = np.zeros(n, dtype=np.int64)
    pairs = []
    choices = rng.integers(0, 1 << 30, size=n)
    for i in range(1, n):
        lo = max(0, i - 3)
        cand = [p for p in range(lo, i) if deg[p] < 4]
        back = lo
        while not cand:                      # window full: widen it backwards
            back -= 1
            if deg[back] < 4:
                cand = [back]
        p = cand[int(choices[i]) % len(cand)]
        pairs.append((p, i))
        deg[p] += 1
        deg[i] += 1
    n_ring = int(round(0.018 * n))
    if n_ring and n > 6:
        starts = rng.choice(n - 5, size=min(n_ring, n - 5), replace=False)
        for s in np.sort(starts):
            pairs.append((int(s), int(s) + 5))
    e = np.empty((2, 2 * len(pairs)), dtype=np.int64)
    if pairs:
        a = np.asarray(pairs, dtype=np.int64)
        e[0, 0::2], e[1, 0::2] = a[:, 0], a[:, 1]
        e[0, 1::2], e[1, 1::2] = a[:, 1], a[:, 0]
    return e


def peptides_graphs(num_graphs: int, seed: int = 1234, task: str = "func",
                    fixed_nodes: Optional[int] = None) -> List[Data]:
    """Peptides-shaped graphs: x int64 [n,9] (cast with .float() by the caller, train.py:79)."""
    rng = np.random.default_rng(seed)
    out: List[Data] = []
    for _ in range(num_graphs):
        if fixed_nodes is not None:
            n = int(fixed_nodes)
        else:
            n = int(np.clip(np.rint(rng.normal(150.94, 60.0)), 8, 444))
        ei = _peptide_edges(n, rng)
        x = np.stack([rng.integers(0, r, size=n) for r in ATOM_FEATURE_RANGES], axis=1).astype(np.int64)
        if task == "func":
            y = (rng.random((1, 10)) < 0.15).astype(np.float32)
        else:
            y = rng.normal(0.0, 1.0, size=(1, 11)).astype(np.float32)
        out.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei), y=torch.from_numpy(y)))
    return out


def _peptide_size(rng: np.random.Generator) -> int:
    return int(np.clip(np.rint(rng.normal(150.94, 60.0)), 8, 444))


def peptides_graph_size(seed: int, index: int) -> int:
    """Node count of `peptides_graph(seed, index)` without building it (data-parallel ranks balance a global batch
    by size and then generate only their own graphs)."""
    return _peptide_size(np.random.default_rng([int(seed), int(index)]))


def peptides_graph(seed: int, index: int, task: str = "func") -> Data:
    """Graph `index` of the global batch `seed`: same distribution as `peptides_graphs`, but every graph has its own
    random stream, so any subset can be generated independently and identically on every rank."""
    rng = np.random.default_rng([int(seed), int(index)])
    n = _peptide_size(rng)
    ei = _peptide_edges(n, rng)
    x = np.stack([rng.integers(0, r, size=n) for r in ATOM_FEATURE_RANGES], axis=1).astype(np.int64)
    if task == "func":
        y = (rng.random((1, 10)) < 0.15).astype(np.float32)
    else:
        y = rng.normal(0.0, 1.0, size=(1, 11)).astype(np.float32)
    return Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei), y=torch.from_numpy(y))


def vocsp_graphs(num_graphs: int, seed: int = 1238, num_classes: int = 21) -> List[Data]:
    """PascalVOC-SP-shaped graphs: n in [395,500], ~2.83 links/node within index distance 16, symmetrised."""
    rng = np.random.default_rng(seed)
    out: List[Data] = []
    for _ in range(num_graphs):
        n = int(rng.integers(395, 501))
        k = rng.poisson(2.83, size=n).clip(0, 8)
        src = np.repeat(np.arange(n), k)
        off = rng.integers(1, 17, size=src.size) * rng.choice([-1, 1], size=src.size)
        dst = src + off
        ok = (dst >= 0) & (dst < n)
        src, dst = src[ok], dst[ok]
        und = np.unique(np.stack([np.minimum(src, dst), np.maximum(src, dst)], 1), axis=0)
        order = rng.permutation(und.shape[0])
        und = und[order]
        e = np.empty((2, 2 * und.shape[0]), dtype=np.int64)
        e[0, 0::2], e[1, 0::2] = und[:, 0], und[:, 1]
        e[0, 1::2], e[1, 1::2] = und[:, 1], und[:, 0]
        x = rng.normal(0.0, 1.0, size=(n, 14)).astype(np.float32)
        y = rng.integers(0, num_classes, size=n).astype(np.int64)
        out.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(e), y=torch.from_numpy(y)))
    return out


def erdos_renyi_graphs(num_graphs: int, nodes_per_graph: int, avg_degree: float, num_features: int,
                       seed: int = 1239) -> List[Data]:
    """Directed ER graphs for the SpMM avg-degree sweep (config #5)."""
    rng = np.random.default_rng(seed)
    out: List[Data] = []
    for _ in range(num_graphs):
        n = nodes_per_graph
        m = int(round(avg_degree * n))
        src = rng.integers(0, n, size=m)
        dst = rng.integers(0, n, size=m)
        e = np.stack([src, dst]).astype(np.int64)
        x = rng.normal(0.0, 1.0, size=(n, num_features)).astype(np.float32)
        out.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(e)))
    return out


def peptides_batch(num_graphs: int, seed: int = 1234, task: str = "func",
                   fixed_nodes: Optional[int] = None) -> Batch:
    return Batch.from_data_list(peptides_graphs(num_graphs, seed, task, fixed_nodes))
