"""numpy restatement of the cluster -> virtual-node construction (loader/hetero_data.py:42-87).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Integer outputs are compared bit-exactly with K7.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
from torch import Tensor


def cluster_argmax(s_soft: Tensor) -> np.ndarray:
    """train_clustering.py:68 -- `clust.max(1)[1].cpu().numpy()` (first maximal index)."""
    return s_soft.max(1)[1].cpu().numpy()


def virtual_nodes(x: Tensor, clusters: np.ndarray, num_clusters: int
                  ) -> Tuple[np.ndarray, Tensor, Tensor, Tensor]:
    """-> (remapped cluster ids [n], virtual.x fp32 [U,F], vv edge_index [2,U(U+1)/2], lv edge_index [2,n]).

    hetero_data.py:45-51  remap to 0..U-1 by sorted unique
    hetero_data.py:52-54  bucket node i into slot remapped-1 (Python negative index: cluster 0 -> last slot)
    hetero_data.py:55-59  drop empty slots, np.mean (float64) per slot
    hetero_data.py:66     torch.FloatTensor(...) rounds once to fp32
    hetero_data.py:68-79  vv = [col; row], col = [a]*(U-a), row = range(U-a), for a in range(U)
    hetero_data.py:80-86  lv = (i, remapped_i)
    """
    uniq = np.unique(clusters)
    lut = {int(v): i for i, v in enumerate(uniq)}
    remapped = np.asarray([lut[int(v)] for v in clusters], dtype=np.int64)
    buckets: List[List[list]] = [[] for _ in range(num_clusters)]
    rows = x.tolist()
    for i, c in enumerate(remapped):
        buckets[int(c) - 1].append(rows[i])
    buckets = [b for b in buckets if len(b) != 0]
    means = np.array([np.mean(b, axis=0) for b in buckets])
    U = len(means)
    virt_x = torch.FloatTensor(means)
    col = np.concatenate([[a] * (U - a) for a in range(U)])
    row = np.concatenate([list(range(U - a)) for a in range(U)])
    vv = torch.LongTensor(np.stack([col, row]).astype(np.int64))
    lv = torch.LongTensor(np.stack([np.arange(len(remapped)), remapped]).astype(np.int64))
    return remapped, virt_x, vv, lv
