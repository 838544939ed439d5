"""Graph containers and the batch/ptr collate convention the hot path consumes.

Host-side mirror of the subset of torch_geometric.data / torch_geometric.loader
the reference touches ("Batch / ptr convention"):

  * `Data`        -- mpnn.py:6,49-50 ; train_clustering.py:36-47 (x, edge_index, edge_weight, y, num_nodes)
  * `HeteroData`  -- loader/hetero_data.py:62-86 (h["local"].x, h["local","to","virtual"].edge_index, ...)
  * `Batch`       -- hscn.py:5,111 ; train.py:75-81 (batch.x_dict, batch.edge_index_dict, batch["local"].batch)
  * `DataLoader`  -- loader/loader.py:48-60

Collate rules (PyG `Batch.from_data_list`): node-level
tensors are concatenated on dim 0, `edge_index` on dim 1 with the cumulative
node count added (row 0 by source-type offset, row 1 by destination-type
offset for heterogeneous graphs), and `batch` (int64, non-decreasing) plus
`ptr` (int64, B+1) are emitted per node type.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

NodeType = str
EdgeType = Tuple[str, str, str]


def _is_index_key(key: str) -> bool:
    return "index" in key or key == "face"


class _Store:
    """Attribute bag for one node type / edge type / homogeneous graph."""

    def __init__(self, **kwargs: Any):
        object.__setattr__(self, "_d", dict(kwargs))

    def __getattr__(self, key: str) -> Any:
        d = object.__getattribute__(self, "_d")
        if key in d:
            return d[key]
        raise AttributeError(key)

    def __setattr__(self, key: str, value: Any) -> None:
        self._d[key] = value

    def __getitem__(self, key: str) -> Any:
        return self._d[key]

    def __setitem__(self, key: str, value: Any) -> None:
        self._d[key] = value

    def __contains__(self, key: str) -> bool:
        return key in self._d

    def keys(self) -> List[str]:
        return list(self._d.keys())

    def items(self):
        return self._d.items()

    def get(self, key: str, default: Any = None) -> Any:
        return self._d.get(key, default)

    def _apply(self, fn) -> "_Store":
        out = type(self)()
        for k, v in self._d.items():
            out._d[k] = fn(v) if isinstance(v, Tensor) else v
        return out

    @property
    def num_nodes(self) -> Optional[int]:
        d = self._d
        if "num_nodes" in d:
            return int(d["num_nodes"])
        if isinstance(d.get("x"), Tensor):
            return d["x"].size(0)
        if isinstance(d.get("batch"), Tensor):
            return d["batch"].size(0)
        return None


class Data:
    """Homogeneous graph.  Missing standard attributes read as None (PyG behaviour
    relied on at train_clustering.py:39 where `data.edge_weight` is absent)."""

    _STANDARD = ("x", "edge_index", "edge_attr", "edge_weight", "y", "pos", "batch", "ptr")

    def __init__(self, x: Optional[Tensor] = None, edge_index: Optional[Tensor] = None,
                 edge_attr: Optional[Tensor] = None, y: Optional[Tensor] = None, **kwargs: Any):
        object.__setattr__(self, "_store", _Store())
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, **kwargs).items():
            if v is not None:
                self._store[k] = v

    def __getattr__(self, key: str) -> Any:
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        if key in Data._STANDARD:
            return None
        raise AttributeError(f"'{type(self).__name__}' has no attribute '{key}'")

    def __setattr__(self, key: str, value: Any) -> None:
        if value is None and key in self._store:
            del self._store._d[key]
        elif value is not None:
            self._store[key] = value

    def __getitem__(self, key: str) -> Any:
        return self._store[key]

    def __setitem__(self, key: str, value: Any) -> None:
        self._store[key] = value

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def keys(self) -> List[str]:
        return self._store.keys()

    @property
    def num_nodes(self) -> int:
        n = self._store.num_nodes
        if n is not None:
            return n
        ei = self._store.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @property
    def num_edges(self) -> int:
        ei = self._store.get("edge_index")
        return 0 if ei is None else ei.size(1)

    @property
    def num_features(self) -> int:
        x = self._store.get("x")
        return 0 if x is None else (1 if x.dim() == 1 else x.size(-1))

    num_node_features = num_features

    def _wrap(self, store: _Store) -> "Data":
        out = type(self).__new__(type(self))
        object.__setattr__(out, "_store", store)
        return out

    def to(self, device, non_blocking: bool = False) -> "Data":
        return self._wrap(self._store._apply(lambda t: t.to(device, non_blocking=non_blocking)))

    def pin_memory(self) -> "Data":
        return self._wrap(self._store._apply(lambda t: t.pin_memory()))

    def clone(self) -> "Data":
        return self._wrap(self._store._apply(lambda t: t.clone()))

    def __repr__(self) -> str:
        parts = [f"{k}={list(v.shape) if isinstance(v, Tensor) else v}" for k, v in self._store.items()]
        return f"{type(self).__name__}({', '.join(parts)})"


class HeteroData:
    """Heterogeneous graph keyed by node type (str) and edge type (src, rel, dst)."""

    def __init__(self) -> None:
        object.__setattr__(self, "_nodes", {})   # insertion order == PyG's
        object.__setattr__(self, "_edges", {})

    def __getitem__(self, key: Union[NodeType, EdgeType]) -> _Store:
        if isinstance(key, tuple):
            if len(key) == 2:
                key = (key[0], "to", key[1])
            return self._edges.setdefault(tuple(key), _Store())
        return self._nodes.setdefault(key, _Store())

    def __getattr__(self, key: str) -> Any:
        if key.endswith("_dict") and not key.startswith("_"):
            name = key[:-5]
            out: Dict[Any, Any] = {}
            for t, s in self._nodes.items():
                if name in s:
                    out[t] = s[name]
            for t, s in self._edges.items():
                if name in s:
                    out[t] = s[name]
            return out
        raise AttributeError(key)

    @property
    def node_types(self) -> List[NodeType]:
        return list(self._nodes.keys())

    @property
    def edge_types(self) -> List[EdgeType]:
        return list(self._edges.keys())

    @property
    def num_nodes(self) -> int:
        return sum(int(s.num_nodes or 0) for s in self._nodes.values())

    def _map(self, fn) -> "HeteroData":
        out = type(self)()
        for t, s in self._nodes.items():
            out._nodes[t] = s._apply(fn)
        for t, s in self._edges.items():
            out._edges[t] = s._apply(fn)
        for k, v in self.__dict__.items():
            if k not in ("_nodes", "_edges"):
                out.__dict__[k] = v
        return out

    def to(self, device, non_blocking: bool = False) -> "HeteroData":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "HeteroData":
        return self._map(lambda t: t.pin_memory())

    def __repr__(self) -> str:
        n = {t: {k: list(v.shape) for k, v in s.items() if isinstance(v, Tensor)} for t, s in self._nodes.items()}
        e = {t: {k: list(v.shape) for k, v in s.items() if isinstance(v, Tensor)} for t, s in self._edges.items()}
        return f"{type(self).__name__}(nodes={n}, edges={e})"


class Batch(Data):
    """Collated homogeneous mini-batch; `Batch.from_data_list` also accepts HeteroData."""

    @staticmethod
    def from_data_list(data_list: Sequence[Union[Data, HeteroData]]):
        if len(data_list) and isinstance(data_list[0], HeteroData):
            return HeteroBatch.from_data_list(data_list)
        out = Batch()
        keys: List[str] = []
        for d in data_list:
            for k in d.keys():
                if k not in keys:
                    keys.append(k)
        counts = torch.tensor([d.num_nodes for d in data_list], dtype=torch.long)
        ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
        for k in keys:
            vals = [d[k] for d in data_list if k in d]
            if k in ("num_nodes", "batch", "ptr"):
                continue
            if not isinstance(vals[0], Tensor):
                out[k] = vals
            elif _is_index_key(k):
                out[k] = torch.cat([v + int(ptr[i]) for i, v in enumerate(vals)], dim=-1)
            else:
                vals = [v.unsqueeze(0) if v.dim() == 0 else v for v in vals]
                out[k] = torch.cat(vals, dim=0)
        out["batch"] = torch.repeat_interleave(torch.arange(len(data_list), dtype=torch.long), counts)
        out["ptr"] = ptr
        out["num_graphs"] = len(data_list)
        # collate-time facts the device side would otherwise need a host sync for (structure.py): the largest graph,
        # and -- edges are concatenated graph-major with node offsets added, so the batch is block diagonal -- the
        # largest per-graph edge count, which sizes the per-graph CSR kernel (K1 fast path)
        out["max_nodes_per_graph"] = int(counts.max()) if len(data_list) else 0
        if len(data_list) and all("edge_index" in d for d in data_list) and out["edge_index"].numel():
            # verified here, once, on the host: every edge stays inside its graph (a Data object with an out-of-range
            # index would break that); batches that fail keep the general radix path
            from .structure import edge_blocks_from_batch
            try:
                blocks = edge_blocks_from_batch(out["edge_index"], out["batch"], len(data_list))
            except (IndexError, RuntimeError):
                blocks = None
            if blocks is not None:
                out["max_edges_per_graph"] = blocks[1]
        return out

    def to(self, device, non_blocking: bool = False) -> "Batch":
        out = super().to(device, non_blocking=non_blocking)
        out.register_structures()
        return out

    def register_structures(self) -> bool:
        """Tell the structure cache what collate knows about this (CUDA-resident) batch: `ptr` (no sortedness
        check, no `batch.max()` sync) and that the edge list is block diagonal and graph-major (K1 fast path).
        `to(device)` calls it; call it again after `structure_cache().clear()`."""
        ei, b, ptr = self.edge_index, self.batch, self.ptr
        if not (ei is not None and ei.is_cuda and b is not None and ptr is not None and "max_edges_per_graph" in self
                and "num_graphs" in self):
            return False
        from .structure import structure_cache
        cache = structure_cache()
        seg = cache.register_segments(b, ptr, int(self["num_graphs"]), int(self["max_nodes_per_graph"]))
        if int(self["max_edges_per_graph"]) > 0 and int(self["max_nodes_per_graph"]) > 0:
            cache.register_blocks(ei, seg.ptr, int(self["num_graphs"]), int(self["max_nodes_per_graph"]),
                                  int(self["max_edges_per_graph"]))
        return True

    def to_data_list(self) -> List[Data]:
        ptr, B = self.ptr, int(self.num_graphs)
        ei = self.edge_index
        egraph = self.batch[ei[0]] if ei is not None else None
        out = []
        for g in range(B):
            lo, hi = int(ptr[g]), int(ptr[g + 1])
            d = Data()
            for k in self.keys():
                v = self[k]
                if k in ("batch", "ptr", "num_graphs") or not isinstance(v, Tensor):
                    continue
                if _is_index_key(k):
                    d[k] = v[:, egraph == g] - lo
                elif v.size(0) == self.batch.numel():
                    d[k] = v[lo:hi]
                elif egraph is not None and v.size(0) == ei.size(1):
                    d[k] = v[egraph == g]
                elif v.size(0) == B:
                    d[k] = v[g:g + 1]
            out.append(d)
        return out


class HeteroBatch(HeteroData):
    """Collated heterogeneous mini-batch (train.py:75-77 consumes x_dict / edge_index_dict / ['local'].batch)."""

    @staticmethod
    def from_data_list(data_list: Sequence[HeteroData]) -> "HeteroBatch":
        out = HeteroBatch()
        B = len(data_list)
        ptrs: Dict[NodeType, Tensor] = {}
        for t in data_list[0].node_types:
            counts = torch.tensor([int(d[t].num_nodes or 0) for d in data_list], dtype=torch.long)
            ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
            ptrs[t] = ptr
            store = out[t]
            for k in data_list[0][t].keys():
                vals = [d[t][k] for d in data_list]
                if isinstance(vals[0], Tensor):
                    vals = [v.unsqueeze(0) if v.dim() == 0 else v for v in vals]
                    store[k] = torch.cat(vals, dim=0)
                elif k != "num_nodes":
                    store[k] = vals
            store["batch"] = torch.repeat_interleave(torch.arange(B, dtype=torch.long), counts)
            store["ptr"] = ptr
        for et in data_list[0].edge_types:
            src, _, dst = et
            store = out[et]
            for k in data_list[0][et].keys():
                vals = [d[et][k] for d in data_list]
                if isinstance(vals[0], Tensor) and _is_index_key(k):
                    shifted = []
                    for i, v in enumerate(vals):
                        off = torch.tensor([[int(ptrs[src][i])], [int(ptrs[dst][i])]], dtype=v.dtype)
                        shifted.append(v + off)
                    store[k] = torch.cat(shifted, dim=-1)
                elif isinstance(vals[0], Tensor):
                    store[k] = torch.cat(vals, dim=0)
                else:
                    store[k] = vals
        out.__dict__["num_graphs"] = B
        return out


class DataLoader(torch.utils.data.DataLoader):
    """torch_geometric.loader.DataLoader mirror (loader.py:48-60): collates with Batch.from_data_list."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, **kwargs: Any):
        kwargs.pop("collate_fn", None)
        kwargs.pop("follow_batch", None)
        kwargs.pop("exclude_keys", None)
        if kwargs.get("num_workers", 0) == 0:
            kwargs.pop("persistent_workers", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=Batch.from_data_list, **kwargs)
