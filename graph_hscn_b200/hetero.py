"""On-device cluster assignment -> heterogeneous (local + virtual node) batch  (K7).

Replaces, for a whole `batch`/`ptr` mini-batch at once and without per-node Python:
    train/train_clustering.py:65-69   clusters = softmax(s).max(1)[1]
    loader/hetero_data.py:42-87       HeteroData with virtual cluster nodes
    loader/hetero_data.py:91-104      collate of those HeteroData into a batch
Two layouts:
  * compact (default): exactly PyG's collated layout -- V = sum_g U_g virtual nodes, U_g(U_g+1)/2 v->v
    edges per graph.  Costs one device->host read of two integers (V, E_vv).
  * padded: K virtual slots and K(K+1)/2 v->v edge slots per graph, unused slots zero / -1 (ignored by
    the CSR build).  Shapes are static, so the whole step can be captured in a CUDA graph.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops
from ._lib import lib
from .data import HeteroBatch
from .structure import CSR, _p, _require_cuda, _stream, structure_cache

LL = ("local", "to", "local")
VV = ("virtual", "to", "virtual")
LV = ("local", "to", "virtual")


_PADDED_LAYOUT: dict = {}


def _padded_layout(B: int, K: int, dev) -> tuple:
    """`batch` / `ptr` of the padded virtual layout (K slots per graph): constants of (B, K), built once per device
    instead of five small launches per step."""
    key = (B, K, dev.index if dev.index is not None else torch.cuda.current_device())
    hit = _PADDED_LAYOUT.get(key)
    if hit is None:
        if torch.cuda.is_current_stream_capturing():       # first use inside a capture: do not cache pool memory
            return torch.arange(B, device=dev).repeat_interleave(K), torch.arange(B + 1, device=dev) * K
        hit = _PADDED_LAYOUT[key] = (torch.arange(B, device=dev).repeat_interleave(K),
                                     torch.arange(B + 1, device=dev) * K)
    return hit


def assign_clusters(s_soft: Tensor) -> Tensor:
    """First-max cluster id per node (int32 [N]); bit-exact with `Tensor.max(1)[1]`."""
    _require_cuda(s_soft)
    return torch.ops.ghscn.cluster_argmax(s_soft.float())


def build_hetero_batch(x_raw: Tensor, edge_index: Tensor, batch: Tensor, clusters: Tensor, num_clusters: int,
                       y: Optional[Tensor] = None, padded: bool = False, num_graphs: Optional[int] = None,
                       x_float: Optional[Tensor] = None) -> HeteroBatch:
    """x_raw: int64 (OGB atoms) or float32 [N,F]; clusters: int32 [N] from `assign_clusters`."""
    _require_cuda(x_raw, edge_index, batch, clusters)
    seg = structure_cache().segments(batch, num_graphs)
    ptr, B, N, K = seg.ptr, seg.num_segments, x_raw.size(0), int(num_clusters)
    dev = x_raw.device
    L, st = lib(), _stream()
    remap, num_virtual, vx_pad = torch.ops.ghscn.virtual_build(clusters.int(), ptr, x_raw, K)
    F = x_raw.size(1)
    lv = torch.empty((2, N), dtype=torch.int64, device=dev)
    out = HeteroBatch()
    voff = torch.empty(B + 1, dtype=torch.int32, device=dev)
    eoff = torch.empty(B + 1, dtype=torch.int32, device=dev)
    L.call("ghscn_virtual_offsets", _p(num_virtual), B, _p(voff), _p(eoff), st)
    if padded:
        V = B * K
        vv_cap = B * (K * (K + 1) // 2)
        vv = torch.empty((2, vv_cap), dtype=torch.int64, device=dev)
        L.call("ghscn_virtual_edges", _p(remap), _p(ptr), _p(num_virtual), None, B, N, K, _p(lv), _p(vv), vv_cap,
               None, st)
        virt_x = vx_pad
        virt_batch, virt_ptr = _padded_layout(B, K, dev)
    else:
        V, E_vv = torch.stack([voff[-1], eoff[-1]]).tolist()      # the one host sync of the compact layout
        vv_cap = E_vv
        vv = torch.empty((2, E_vv), dtype=torch.int64, device=dev)
        L.call("ghscn_virtual_edges", _p(remap), _p(ptr), _p(num_virtual), _p(voff), B, N, K, _p(lv), _p(vv), E_vv,
               _p(eoff), st)
        virt_x = torch.empty((V, F), dtype=torch.float32, device=dev)
        virt_batch = torch.empty(V, dtype=torch.int64, device=dev)
        L.call("ghscn_virtual_compact", _p(vx_pad), _p(num_virtual), _p(voff), B, K, F, _p(virt_x), _p(virt_batch),
               st)
        virt_ptr = voff.long()
    local = out["local"]
    if x_float is not None:          # already cast by the caller (on ITS stream: see train.GraphHSCNStep)
        local.x = x_float
    else:
        local.x = ops.cast_i64_f32(x_raw) if x_raw.dtype == torch.int64 else x_raw
    if y is not None:
        local.y = y
    local.batch, local.ptr = batch, ptr.long()
    virt = out["virtual"]
    virt.x, virt.batch, virt.ptr = virt_x, virt_batch, virt_ptr
    out[LL].edge_index = edge_index      # insertion order ll, vv, lv == hetero_data.py:67,77,84
    out[VV].edge_index = vv
    out[LV].edge_index = lv
    # K7 also emits both CSR orientations of the two virtual relations, so the layers never sort them
    i32 = dict(dtype=torch.int32, device=dev)
    lvd = CSR(torch.empty(V + 1, **i32), torch.empty(N, **i32), torch.empty(N, **i32), V, N, N)
    lvs = CSR(torch.empty(N + 1, **i32), torch.empty(N, **i32), torch.empty(N, **i32), N, N, N)
    vvd = CSR(torch.empty(V + 1, **i32), torch.empty(vv_cap, **i32), torch.empty(vv_cap, **i32), V, vv_cap, vv_cap)
    vvs = CSR(torch.empty(V + 1, **i32), torch.empty(vv_cap, **i32), torch.empty(vv_cap, **i32), V, vv_cap, vv_cap)
    L.call("ghscn_virtual_csr", _p(remap), _p(ptr), _p(num_virtual), _p(voff), _p(eoff), B, K, int(padded),
           _p(lvd.rowptr), _p(lvd.col), _p(lvd.perm), _p(lvs.rowptr), _p(lvs.col), _p(lvs.perm),
           _p(vvd.rowptr), _p(vvd.col), _p(vvd.perm), _p(vvs.rowptr), _p(vvs.col), _p(vvs.perm), st)
    cache = structure_cache()
    cache.register_graph(lv, N, V, by_dst=lvd, by_src=lvs)
    cache.register_graph(vv, V, V, by_dst=vvd, by_src=vvs)
    out.__dict__["num_graphs"] = B
    out.__dict__["cluster"] = remap
    out.__dict__["num_virtual"] = num_virtual
    return out
