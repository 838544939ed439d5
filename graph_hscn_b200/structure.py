"""Batch graph structures in HBM: CSR built once per batch, reused by every layer, fwd and bwd.

The reference recomputes `gcn_norm` and works on raw COO in every GCNConv call of every layer
(`cached=False`, model/hscn.py:124; model/mpnn.py:29-32).  Here a `GraphStructure` is derived once per
`edge_index` tensor and memoised by tensor identity (data_ptr, version, shape), so the five layers of
the MPNN and the backward pass all share one stable-sorted CSR (by destination), its transpose (by
source) and the per-slot normalised weights.

HBM layout per relation (int32 indices, fp32 weights), M = E (+ N when self loops are appended):
    rowptr  [R+1]   col [M]   perm [M]          rows = destination nodes  (forward aggregation)
    rowptr_t[C+1]   col_t[M]  perm_t[M]         rows = source nodes       (backward / A rows of MinCUT)
    w [M], w_t [M]                              per-slot weights in each orientation
"""
from __future__ import annotations

import threading
from collections import OrderedDict
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from ._lib import lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("graph_hscn_b200 kernels are CUDA-only (sm_100a); got a CPU tensor. "
                               "There is no CPU fallback in the product path.")


# ---------------------------------------------------------------------------------------------
# hints: values PyG obtains with a device->host sync (`batch.max()+1`, max nodes per graph).  Inside a
# CUDA-graph capture no sync is possible, so callers provide them.
# ---------------------------------------------------------------------------------------------
_tls = threading.local()


def current_hints() -> Dict[str, int]:
    return getattr(_tls, "hints", {})


@contextmanager
def structure_hints(**kwargs: int):
    """e.g. structure_hints(num_graphs=128, max_nodes_per_graph=444, batch_sorted=1)."""
    old = current_hints()
    _tls.hints = {**old, **kwargs}
    try:
        yield
    finally:
        _tls.hints = old


def _capturing() -> bool:
    return torch.cuda.is_current_stream_capturing()


@dataclass
class CSR:
    rowptr: Tensor
    col: Tensor
    perm: Tensor
    num_rows: int
    num_edges: int     # E (original edges)
    num_items: int     # M = E (+ num_rows with self loops); slots >= rowptr[-1] are unused


def build_csr(key: Tensor, other: Tensor, num_rows: int, add_self_loops: bool = False) -> CSR:
    """Stable sort of the edges by `key` (K1).  Bit-exact with torch.sort(key, stable=True)."""
    _require_cuda(key, other)
    assert key.dtype == torch.int64 and other.dtype == torch.int64
    key, other = key.contiguous(), other.contiguous()
    E = key.numel()
    M = E + (num_rows if add_self_loops else 0)
    dev = key.device
    rowptr = torch.empty(num_rows + 1, dtype=torch.int32, device=dev)
    col = torch.empty(M, dtype=torch.int32, device=dev)
    perm = torch.empty(M, dtype=torch.int32, device=dev)
    L = lib()
    ws_bytes = L.query("ghscn_csr_workspace_bytes", E, num_rows, int(add_self_loops))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    L.call("ghscn_csr_build", _p(key), _p(other), E, num_rows, int(add_self_loops), _p(rowptr), _p(col), _p(perm),
           _p(ws), ws_bytes, _stream())
    return CSR(rowptr, col, perm, num_rows, E, M)


class GraphStructure:
    """Both orientations of one relation's edge set plus cached per-slot weights."""

    MAX_WEIGHT_SETS = 8      # distinct (edge_weight tensor, normalize, improved) combinations kept per structure

    def __init__(self, edge_index: Tensor, num_src: int, num_dst: int, add_self_loops: bool = False):
        _require_cuda(edge_index)
        if add_self_loops and num_src != num_dst:
            raise ValueError("self loops need a square relation")
        self.edge_index = edge_index if edge_index.is_contiguous() else edge_index.contiguous()
        self.num_src, self.num_dst = int(num_src), int(num_dst)
        self.add_self_loops = bool(add_self_loops)
        self._by_dst: Optional[CSR] = None
        self._by_src: Optional[CSR] = None
        self._weights: "OrderedDict[tuple, dict]" = OrderedDict()    # small LRU: entries pin w / w_t / dis
        self._slot_map_t: Optional[Tensor] = None
        self._keep: list = []
        self._parent: Optional["GraphStructure"] = None   # set by StructureCache.alias
        self._plain: Optional["GraphStructure"] = None    # loop-free edge list: derive instead of sorting again
        self._blocks: Optional["EdgeBlocks"] = None       # collated batch: per-graph counting sort, no radix passes

    @property
    def num_edges(self) -> int:
        return self.edge_index.size(1)

    @staticmethod
    def _as_plain(c: CSR) -> CSR:
        return CSR(c.rowptr, c.col, c.perm, c.num_rows, c.num_items, c.num_items)

    @staticmethod
    def _with_loops(c: CSR) -> CSR:
        """K1 shortcut: CSR(edges + appended loops) from CSR(edges) when no edge is a self loop."""
        R, E = c.num_rows, c.num_edges
        dev = c.rowptr.device
        rowptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
        col = torch.empty(E + R, dtype=torch.int32, device=dev)
        perm = torch.empty(E + R, dtype=torch.int32, device=dev)
        lib().call("ghscn_csr_add_loops", _p(c.rowptr), _p(c.col), _p(c.perm), R, E, _p(rowptr), _p(col), _p(perm),
                   _stream())
        return CSR(rowptr, col, perm, R, E, E + R)

    def _build_blocked(self) -> None:
        """K1 fast path: both orientations in ONE launch, one CTA per graph (ghscn_csr_build_blocked)."""
        b = self._blocks
        E, N, dev = self.num_edges, self.num_dst, self.edge_index.device
        rp, ep = (N + 1 + 63) // 64 * 64, (E + 63) // 64 * 64         # 256-byte aligned sub-arrays of one buffer
        out = torch.empty(2 * rp + 4 * ep, dtype=torch.int32, device=dev)
        rp_d, rp_s = out[:N + 1], out[rp:rp + N + 1]
        col_d, perm_d, col_s, perm_s = (out[2 * rp + i * ep:2 * rp + i * ep + E] for i in range(4))
        lib().call("ghscn_csr_build_blocked", _p(self.edge_index[0]), _p(self.edge_index[1]), E, _p(b.ptr),
                   b.num_graphs, N, b.max_nodes, b.max_edges, _p(rp_d), _p(col_d), _p(perm_d), _p(rp_s), _p(col_s),
                   _p(perm_s), _p(b.status), _stream())
        self._by_dst = CSR(rp_d, col_d, perm_d, N, E, E)
        self._by_src = CSR(rp_s, col_s, perm_s, N, E, E)

    def _blocked_ok(self) -> bool:
        b = self._blocks
        if b is None or self.add_self_loops or self.num_src != self.num_dst or self.num_edges == 0:
            return False
        return lib().query("ghscn_csr_blocked_smem_bytes", b.max_nodes, b.max_edges) <= 200 * 1024

    @property
    def by_dst(self) -> CSR:
        if self._by_dst is None and self._blocked_ok():
            self._build_blocked()
        if self._by_dst is None and self._parent is not None:
            self._by_dst = self._as_plain(self._parent.by_dst)
        if self._by_dst is None and self._plain is not None:
            self._by_dst = self._with_loops(self._plain.by_dst)
        if self._by_dst is None:
            self._by_dst = build_csr(self.edge_index[1], self.edge_index[0], self.num_dst, self.add_self_loops)
        return self._by_dst

    @property
    def by_src(self) -> CSR:
        if self._by_src is None and self._blocked_ok():
            self._build_blocked()
        if self._by_src is None and self._parent is not None:
            self._by_src = self._as_plain(self._parent.by_src)
        if self._by_src is None and self._plain is not None:
            self._by_src = self._with_loops(self._plain.by_src)
        if self._by_src is None:
            self._by_src = build_csr(self.edge_index[0], self.edge_index[1], self.num_src, self.add_self_loops)
        return self._by_src

    @property
    def slot_map_t(self) -> Tensor:
        """map_t[t] = by_dst slot of the edge stored in by_src slot t."""
        if self._slot_map_t is None:
            d, s = self.by_dst, self.by_src
            scratch = torch.empty(d.num_items, dtype=torch.int32, device=d.perm.device)
            out = torch.empty(d.num_items, dtype=torch.int32, device=d.perm.device)
            lib().call("ghscn_slot_map", _p(d.perm), _p(s.perm), d.num_items, d.num_items, _p(scratch), _p(out),
                       _stream())
            self._slot_map_t = out
        return self._slot_map_t

    # -- weights --------------------------------------------------------------------------------
    def _loop_weight(self, edge_weight: Optional[Tensor], fill: float) -> Optional[Tensor]:
        if not self.add_self_loops or (edge_weight is None and fill == 1.0):
            return None
        dev = self.edge_index.device
        scratch = torch.empty(self.num_dst, dtype=torch.int32, device=dev)
        lw = torch.empty(self.num_dst, dtype=torch.float32, device=dev)
        lib().call("ghscn_loop_weights", _p(self.edge_index[0]), _p(self.edge_index[1]), _p(edge_weight),
                   self.num_edges, self.num_dst, float(fill), _p(scratch), _p(lw), _stream())
        return lw

    def weights(self, edge_weight: Optional[Tensor], normalize: bool, improved: bool = False,
                need_transpose: bool = True) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
        """(w, w_t, dis): per-slot weights for by_dst / by_src; None means unit weights."""
        if edge_weight is not None:
            _require_cuda(edge_weight)
            edge_weight = edge_weight.detach().contiguous().float()
        key = (None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version), normalize, improved)
        L, st = lib(), _stream()
        dev = self.edge_index.device
        fill = 2.0 if improved else 1.0
        ent = self._weights.get(key)
        if ent is None:
            if not normalize and edge_weight is None and not (self.add_self_loops and fill != 1.0):
                ent = dict(w=None, w_t=None, dis=None, lw=None, ew=None, unit=True)
            else:
                d = self.by_dst
                lw = self._loop_weight(edge_weight, fill)
                dis = None
                if normalize:
                    dis = torch.empty(self.num_dst, dtype=torch.float32, device=dev)
                    L.call("ghscn_gcn_deg_inv_sqrt", _p(d.rowptr), _p(d.perm), _p(edge_weight), _p(lw),
                           d.num_edges, d.num_rows, _p(dis), st)
                w = torch.empty(d.num_items, dtype=torch.float32, device=dev)
                L.call("ghscn_edge_weights", _p(d.rowptr), _p(d.col), _p(d.perm), _p(edge_weight), _p(lw), _p(dis),
                       d.num_edges, d.num_rows, int(normalize), 1, _p(w), st)
                ent = dict(w=w, w_t=None, dis=dis, lw=lw, ew=edge_weight, unit=False)
            self._weights[key] = ent
            while len(self._weights) > self.MAX_WEIGHT_SETS:
                self._weights.popitem(last=False)
        else:
            self._weights.move_to_end(key)
        if need_transpose and not ent["unit"] and ent["w_t"] is None:
            s = self.by_src
            w_t = torch.empty(s.num_items, dtype=torch.float32, device=dev)
            L.call("ghscn_edge_weights", _p(s.rowptr), _p(s.col), _p(s.perm), _p(ent["ew"]), _p(ent["lw"]),
                   _p(ent["dis"]), s.num_edges, s.num_rows, int(normalize), 0, _p(w_t), st)
            ent["w_t"] = w_t
        return ent["w"], ent["w_t"], ent["dis"]


@dataclass
class Segments:
    """Contiguous (or permuted) row ranges for segment reductions: `batch` / `ptr` convention."""
    ptr: Tensor                    # int32 [B+1]
    perm: Optional[Tensor]         # int32 [N] gather order when the index vector is not sorted
    num_segments: int
    num_rows: int
    max_rows: int = 0              # max rows of any segment (0 = unknown)


@dataclass
class EdgeBlocks:
    """Collate-time knowledge about an edge list: edges are stored graph-major and stay inside their graph."""
    ptr: Tensor                    # int32 [B+1] node ranges
    num_graphs: int
    max_nodes: int                 # max nodes of any graph
    max_edges: int                 # max edges of any graph
    status: Tensor                 # int32 [1], non-zero if a kernel found the promise broken
    index: Optional[Tensor] = None # the edge list the promise was made for: pinned, so its address cannot be recycled


def edge_blocks_from_batch(edge_index: Tensor, batch: Tensor, num_graphs: Optional[int] = None
                           ) -> Optional[Tuple[int, int]]:
    """(max nodes per graph, max edges per graph) if `edge_index` is block diagonal and graph-major w.r.t. the
    sorted `batch` vector, else None.  Reads the result back (host sync): collate time / eager mode only."""
    if edge_index.numel() == 0 or batch.numel() == 0:
        return None
    g_src, g_dst = batch[edge_index[0]], batch[edge_index[1]]
    ok = bool((g_src == g_dst).all()) and bool((g_src[1:] >= g_src[:-1]).all()) and bool((batch[1:] >= batch[:-1]).all())
    if not ok:
        return None
    B = int(num_graphs) if num_graphs is not None else int(batch[-1]) + 1
    return int(torch.bincount(batch, minlength=B).max()), int(torch.bincount(g_src, minlength=B).max())


class StructureCache:
    """Memoises structures by tensor identity; the cached entry pins the index tensor."""

    def __init__(self, capacity: int = 64):
        self.capacity = capacity
        self._graphs: "OrderedDict[tuple, GraphStructure]" = OrderedDict()
        self._segments: "OrderedDict[tuple, Tuple[Tensor, Segments]]" = OrderedDict()
        self._blocks: "OrderedDict[tuple, EdgeBlocks]" = OrderedDict()
        self._status: Dict[int, Tensor] = {}
        self.builds = 0

    def blocked_status(self, device) -> Tensor:
        """Device flag OR-ed by ghscn_csr_build_blocked when its preconditions do not hold (0 = fine)."""
        idx = torch.device(device).index or 0
        if idx not in self._status:
            self._status[idx] = torch.zeros(1, dtype=torch.int32, device=device)
        return self._status[idx]

    def check_blocked_status(self, device) -> None:
        """Host read of the flag (one sync; never inside a capture): raises when a per-graph CSR build found its
        promise broken -- the CSR slots of that batch were left unwritten, nothing computed from them is valid."""
        flag = int(self.blocked_status(device).item())
        if flag:
            self.blocked_status(device).zero_()
            raise RuntimeError(f"ghscn_csr_build_blocked: the block-diagonal / graph-major promise registered for an "
                               f"edge list did not hold (status {flag}); rebuild the batch without register_blocks")

    def register_blocks(self, edge_index: Tensor, ptr: Tensor, num_graphs: int, max_nodes: int, max_edges: int
                        ) -> EdgeBlocks:
        """Promise that `edge_index` is the collated edge list of the graphs delimited by `ptr`: its CSRs are then
        built by the per-graph kernel (one launch) instead of the radix passes."""
        b = EdgeBlocks(ptr if ptr.dtype == torch.int32 else ptr.to(torch.int32), int(num_graphs), int(max_nodes),
                       int(max_edges), self.blocked_status(edge_index.device), edge_index)
        self._blocks[self._key(edge_index)] = b
        while len(self._blocks) > self.capacity:
            self._blocks.popitem(last=False)
        return b

    @staticmethod
    def _key(t: Tensor, *extra) -> tuple:
        return (t.data_ptr(), t._version, tuple(t.shape), t.device.index, *extra)

    def clear(self) -> None:
        self._graphs.clear()
        self._segments.clear()
        self._blocks.clear()

    def graph(self, edge_index: Tensor, num_src: int, num_dst: int, add_self_loops: bool = False) -> GraphStructure:
        key = self._key(edge_index, int(num_src), int(num_dst), bool(add_self_loops))
        st = self._graphs.get(key)
        if st is None:
            st = GraphStructure(edge_index, num_src, num_dst, add_self_loops)
            if not add_self_loops:
                st._blocks = self._blocks.get(self._key(edge_index))     # the entry pins its tensor: same key == same tensor
            if add_self_loops and self._known_loop_free(edge_index):
                st._plain = self.graph(edge_index, num_src, num_dst, False)
            self._graphs[key] = st
            self.builds += 1
            while len(self._graphs) > self.capacity:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        return st

    @staticmethod
    def _known_loop_free(edge_index: Tensor) -> bool:
        """True when the edge list is known to hold no self loop (hint, or one cheap check outside capture)."""
        if current_hints().get("no_self_loops"):
            return True
        if _capturing() or edge_index.numel() == 0:
            return False            # unknown: the general sort-with-loops path needs no host sync
        return not bool((edge_index[0] == edge_index[1]).any().item())

    def register_graph(self, edge_index: Tensor, num_src: int, num_dst: int, by_dst: CSR, by_src: CSR
                       ) -> GraphStructure:
        """Install a relation whose CSRs were produced directly (K7 emits l->v / v->v without sorting)."""
        st = GraphStructure(edge_index, num_src, num_dst, False)
        st._by_dst, st._by_src = by_dst, by_src
        self._graphs[self._key(edge_index, int(num_src), int(num_dst), False)] = st
        return st

    def alias(self, new_index: Tensor, parent: GraphStructure) -> GraphStructure:
        """Register `new_index` (= parent's edges followed by parent's appended self loops, nothing dropped)
        as a plain relation sharing parent's sorted arrays: item ids coincide, so no second sort is needed."""
        n = parent.num_dst
        view = GraphStructure(new_index, n, n, False)
        view._parent = parent
        key = self._key(new_index, n, n, False)
        self._graphs[key] = view
        return view

    def register_segments(self, index: Tensor, ptr: Tensor, num_segments: int, max_rows: int = 0) -> Segments:
        """Prime the cache from collate-time knowledge (Batch.to(device) does this): no sync needed later."""
        seg = Segments(ptr.to(device=index.device, dtype=torch.int32), None, int(num_segments), index.numel(),
                       int(max_rows))
        self._segments[self._key(index)] = (index, seg)
        return seg

    def segments(self, index: Tensor, dim_size: Optional[int] = None) -> Segments:
        """Segments of a PyG `batch`-style index vector (sorted or not)."""
        _require_cuda(index)
        key = self._key(index)
        hit = self._segments.get(key)
        if hit is not None and (dim_size is None or hit[1].num_segments == dim_size):
            return hit[1]
        hints = current_hints()
        n = index.numel()
        if dim_size is None:
            dim_size = hints.get("num_graphs")
        if dim_size is None:
            if _capturing():
                raise RuntimeError("segment count needs a host sync; pass size= or structure_hints(num_graphs=)")
            dim_size = int(index.max().item()) + 1 if n else 0
        if hints.get("batch_sorted"):
            is_sorted = True
        elif _capturing():
            raise RuntimeError("sortedness check needs a host sync; use structure_hints(batch_sorted=1)")
        else:
            is_sorted = bool((index[1:] >= index[:-1]).all().item()) if n > 1 else True
        index = index.contiguous()
        if is_sorted:
            ptr = torch.empty(dim_size + 1, dtype=torch.int32, device=index.device)
            lib().call("ghscn_batch_to_ptr", _p(index), n, dim_size, _p(ptr), _stream())
            seg = Segments(ptr, None, dim_size, n, int(hints.get("max_nodes_per_graph", 0)))
        else:
            csr = build_csr(index, index, dim_size, False)
            seg = Segments(csr.rowptr, csr.perm, dim_size, n, 0)
        self._segments[key] = (index, seg)
        self.builds += 1
        while len(self._segments) > self.capacity:
            self._segments.popitem(last=False)
        return seg


_CACHE = StructureCache()


def structure_cache() -> StructureCache:
    return _CACHE


@contextmanager
def capture_scope():
    """Use around CUDA-graph capture: structures built inside live in the graph's memory pool."""
    _CACHE.clear()
    try:
        yield _CACHE
    finally:
        _CACHE.clear()
