"""Graph-HSCN training step on one B200 (CUDA-graph captured per shape bucket) and its data-parallel form.

One step = one pass of the hot path over one `batch`/`ptr` mini-batch (BASELINE config #2):
    1. SCN stage   -- gcn_norm(add_self_loops) -> GraphConv stack -> logits -> fused MinCUT losses,
                      backward, AdamW step                    (train/train_clustering.py:34-50, batched)
    2. assignment  -- SCN forward with the updated weights, softmax, first-max cluster id
                      (train_clustering.py:57-69) and on-device virtual-node construction
                      (loader/hetero_data.py:42-87)
    3. HSCN stage  -- 3-relation HeteroConv stack, mean readout, 2 linears, loss, backward,
                      [clip_grad_norm], AdamW step every `batch_accumulation` batches (train/train.py:73-95)

Variable-shape batches (train/train.py:73 iterates a DataLoader): a CUDA graph needs static shapes, so every batch is
padded into a *bucket* (`BucketPolicy`): node count rounded up to a multiple of `node_step`, edge count to a multiple
of `edge_step`.  The padding is `D` trailing DUMMY GRAPHS -- real, well-formed graphs (zero features, zero labels,
a chain of edges among the pad nodes) appended after the B real ones -- so every structure kernel sees an ordinary
block-diagonal, graph-major batch of B + D graphs.  The losses look at the first B graphs only (MinCUT runs on
`ptr[:B+1]`; the task loss on `pred[:B]`), so the dummy rows carry exactly zero gradient and no real row ever reads
them.  One CUDA graph is captured per bucket on first use and replayed afterwards (`GraphHSCNStep.load/run`).

Data parallelism: graphs are independent, so ranks take disjoint sets of graphs and exchange nothing
in forward/backward; one NCCL all-reduce of a single flat fp32 gradient buffer per model per optimizer step sums the
gradients, each rank's contribution weighted by its share of the global graph count (the reference has no
distributed code).
"""
from __future__ import annotations

import math
import os

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor
from torch.nn.parameter import UninitializedParameter

from . import hetero, models, ops
from .data import Batch
from .structure import capture_scope, edge_blocks_from_batch, structure_cache, structure_hints


# ---------------------------------------------------------------------------------------------
# data-parallel plumbing
# ---------------------------------------------------------------------------------------------
def shard_range(num_graphs: int, rank: int, world: int, weights: Optional[Sequence[int]] = None) -> Tuple[int, int]:
    """Contiguous graph range [lo, hi) of `rank`.  With `weights` (nodes or edges per graph) the cut points
    balance the summed weight instead of the graph count (SURVEY 8e "balance by sum n_g"); the ranks then hold
    different graph counts and must pass them to `FlatGradients.all_reduce_mean(local_graphs=, global_graphs=)`."""
    if weights is None:
        base, rem = divmod(num_graphs, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)
    csum = torch.as_tensor(list(weights), dtype=torch.float64).cumsum(0)
    total = float(csum[-1]) if len(csum) else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(torch.searchsorted(csum, torch.tensor(total * r / world, dtype=torch.float64))))
    cuts.append(num_graphs)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def balanced_partition(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Splits graph indices into `world` groups of EQUAL COUNT (up to one) and nearly equal summed size: graphs are
    exchangeable inside a mini-batch (every loss is a mean over graphs), so any equal-count split is a valid
    data-parallel sharding, and equal counts keep the mean of the rank means equal to the batch mean.  Largest graphs
    first, dealt in snake order (0..W-1, W-1..0, ...); each group is returned in ascending index order."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    groups: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        groups[k if rnd % 2 == 0 else world - 1 - k].append(idx)
    return [sorted(g) for g in groups]


class FlatGradients:
    """All gradients of a module as views into ONE flat fp32 buffer: zeroing is one memset and the
    data-parallel exchange is one all-reduce (K8).  Parameters that never receive a gradient
    (the HSCN l->v / v->v branches are dead w.r.t. the loss) are left out on every rank."""

    def __init__(self, module: nn.Module, live: Optional[Sequence[str]] = None):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if live is not None:
            keep = set(live)
            named = [(n, p) for n, p in named if n in keep]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def backward_into(self, loss: Tensor, accumulate: bool = False) -> None:
        """`loss.backward()` for the live parameters, with the gradients written straight into the flat buffer: the
        autograd engine hands the finished gradients over (`torch.autograd.grad`) and ONE multi-tensor copy places them,
        instead of one `grad += g` kernel per parameter interleaved with the backward chain (and no zero fill before).
        Every value is the one `backward()` would have accumulated onto a zeroed buffer; with `accumulate` the
        gradients are added to what the buffer holds (train/train.py:87-89, `batch_accumulation` > 1)."""
        grads = torch.autograd.grad(loss, self.params)
        if accumulate:
            torch._foreach_add_([p.grad for p in self.params], list(grads))
        else:
            self._place(self.params, grads, self.flat)

    def _place(self, params: Sequence[Tensor], grads: Sequence[Tensor], flat: Tensor) -> None:
        """Copies the finished gradients of `params` (consecutive in the flat layout) into `flat`, their slice of it."""
        if flat.is_cuda and all(g.dtype == torch.float32 for g in grads):
            # the parameters' gradient views tile the flat buffer in order: one gather launch (pointers travel as
            # kernel arguments) instead of a multi-tensor copy that runs ~10 thread blocks for 1.1 MB
            import ctypes
            from ._lib import lib
            from .structure import _p, _stream
            gs = [g if g.is_contiguous() else g.contiguous() for g in grads]
            n = len(gs)
            ptrs = (ctypes.c_void_p * n)(*[g.data_ptr() for g in gs])
            sizes = (ctypes.c_int64 * n)(*[g.numel() for g in gs])
            lib().call("ghscn_gather_flat", ptrs, sizes, n, _p(flat), _stream())
        else:
            torch._foreach_copy_([p.grad for p in params], list(grads))

    def leading(self, prefix: str) -> int:
        """Number of leading parameters whose name starts with `prefix` (0 if they are not a prefix of the layout)."""
        k = 0
        while k < len(self.names) and self.names[k].startswith(prefix):
            k += 1
        return 0 if any(n.startswith(prefix) for n in self.names[k:]) else k

    def backward_reduce_overlapped(self, loss: Tensor, boundary: Tensor, split: int, group=None) -> None:
        """`backward_into` + `all_reduce_mean` with the exchange overlapped with the end of the backward pass: the
        gradients of the parameters behind `boundary` (an activation; everything but the first `split` parameters)
        are complete once the backward has reached it, so their all-reduce (99 % of the bytes for the HSCN: every
        h x h layer and the head) runs while the first layer's backward is still computing; only the first layer's
        few kilobytes are exchanged after it.  Same sums as the single all-reduce (the AVG of a slice is the slice
        of the AVG).  NCCL only (the caller falls back to the two separate calls otherwise)."""
        late, early = self.params[split:], self.params[:split]
        n_early = sum(p.numel() for p in early)
        grads = torch.autograd.grad(loss, late + [boundary])
        self._place(late, grads[:-1], self.flat[n_early:])
        work = dist.all_reduce(self.flat[n_early:], op=dist.ReduceOp.AVG, group=group, async_op=True)
        g_early = torch.autograd.grad(boundary, early, grad_outputs=grads[-1])
        self._place(early, g_early, self.flat[:n_early])
        dist.all_reduce(self.flat[:n_early], op=dist.ReduceOp.AVG, group=group)
        work.wait()

    def flatten_parameters(self) -> Tensor:
        """Re-home the live parameters as views of ONE flat fp32 leaf whose .grad is the flat gradient buffer:
        the optimizer step becomes a single elementwise kernel over one tensor (AdamW is elementwise and all
        parameters share one group, so this is the same update as per-parameter AdamW)."""
        flat_p = torch.empty_like(self.flat)
        off = 0
        for p in self.params:
            n = p.numel()
            flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat_p[off:off + n].view_as(p)
            off += n
        flat_p.requires_grad_(True)
        flat_p.grad = self.flat
        self.flat_param = flat_p
        return flat_p

    def all_reduce_mean(self, world: Optional[int] = None, group=None, local_graphs: Optional[int] = None,
                        global_graphs: Optional[int] = None) -> None:
        """Gradient of the GLOBAL batch mean from the per-rank batch-mean gradients.  Every rank's loss is a mean over
        its own graphs, so its gradient is weighted by local_graphs / global_graphs before the SUM all-reduce; with
        equal shards (the default, counts omitted) that weight is 1 / world."""
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = world or dist.get_world_size(group)
        if world == 1:
            return
        if local_graphs is not None and global_graphs:
            self.flat.mul_(float(local_graphs) / float(global_graphs))
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            return
        if self.flat.is_cuda and dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)      # the division happens inside NCCL's kernel
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


class FlatAdamW:
    """torch.optim.AdamW semantics on one flat parameter buffer, one kernel per step (ghscn_adamw_step); the step
    counter lives on the device, so `step()` is CUDA-graph capturable.  `clip_grad_norm(max_norm)` is
    `nn.utils.clip_grad_norm_` (train/train.py:92-93) over the same flat buffer: a fixed-order two-stage sum of
    squares leaves the clip coefficient in a device scalar that the next `step()` multiplies into every gradient as it
    reads it, instead of one norm kernel per parameter, a stack, and one multiply per parameter."""

    def __init__(self, flat_param: Tensor, flat_grad: Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.p, self.g = flat_param, flat_grad
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat_grad)
        self.exp_avg_sq = torch.zeros_like(flat_grad)
        self._state = torch.zeros(4, dtype=torch.float32, device=flat_grad.device)  # steps, lr/(1-b1^t), sqrt(1-b2^t)
        self.step_count = self._state[0:1]
        self.state = {0: {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self._state}}
        self.clip = torch.ones(2, dtype=torch.float32, device=flat_grad.device)      # (total norm, clip coefficient)
        self._clip_ws: Optional[Tensor] = None
        self._use_clip = False

    def clip_grad_norm(self, max_norm: float = 1.0) -> Tensor:
        """-> device tensor [2] = (total 2-norm, coefficient min(1, max_norm / (norm + 1e-6))); applied by `step()`."""
        from ._lib import lib
        from .structure import _p, _stream
        L = lib()
        n = self.g.numel()
        if self._clip_ws is None:
            self._clip_ws = torch.empty(max(L.query("ghscn_grad_clip_workspace_bytes", n), 4), dtype=torch.uint8,
                                        device=self.g.device)
        L.call("ghscn_grad_clip_scale", _p(self.g), n, float(max_norm), _p(self._clip_ws), self._clip_ws.numel(),
               _p(self.clip), _stream())
        self._use_clip = True
        return self.clip

    def step(self) -> None:
        from ._lib import lib
        from .structure import _p, _stream
        scale = self.clip[1:] if self._use_clip else None
        lib().call("ghscn_adamw_step_scaled", _p(self.p), _p(self.g), _p(self.exp_avg), _p(self.exp_avg_sq),
                   self.p.numel(), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                   float(self.wd), _p(self._state), _p(scale), _stream())
        self._use_clip = False

    def zero_grad(self) -> None:
        self.g.zero_()


def live_parameter_names(module: nn.Module, loss: Tensor) -> List[str]:
    """Names of the parameters reachable from `loss` (consistent on all ranks for the same model)."""
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, [p for _, p in named], allow_unused=True, retain_graph=False)
    return [n for (n, _), g in zip(named, grads) if g is not None]


# ---------------------------------------------------------------------------------------------
# shape buckets
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class BucketPolicy:
    """Static-shape buckets for variable-size mini-batches.  `max_*_per_graph` are dataset-level caps on one graph
    (444 nodes / 906 directed edges for the Peptides shape); they bound the dummy graphs too, so the per-graph
    kernels (blocked CSR, MinCUT) keep the shared-memory footprint they have on the real graphs."""
    max_nodes_per_graph: int
    max_edges_per_graph: int
    node_step: int = 256
    edge_step: int = 512
    min_pad_nodes: int = 2          # a dummy graph needs two nodes to carry pad edges without self loops

    @property
    def eff_edge_step(self) -> int:
        return max(1, min(self.edge_step, self.max_edges_per_graph))

    max_pad_degree: int = 4         # pad edges per pad node: keeps the dummy graphs' rows as short as real ones

    @property
    def dummy_graphs(self) -> int:
        worst = 2 * self.node_step + self.min_pad_nodes          # one extra node step when the pad edges need it
        return max(1, math.ceil(worst / max(self.max_nodes_per_graph, 2)))

    def bucket(self, num_nodes: int, num_edges: int) -> Tuple[int, int]:
        """-> (n_cap, e_cap).  The pad edges all land on the pad nodes (a dummy graph cannot borrow nodes from a real
        one), so a bucket with many pad edges and few pad nodes would create rows hundreds of slots long -- one warp of
        the row-parallel kernels would walk them alone (measured: +25 % on the aggregation kernel).  The node capacity
        therefore grows by one more step whenever the pad edges would exceed `max_pad_degree` per pad node."""
        n_cap = math.ceil((num_nodes + self.min_pad_nodes) / self.node_step) * self.node_step
        es = self.eff_edge_step
        e_cap = math.ceil(num_edges / es) * es
        if e_cap - num_edges > self.max_pad_degree * (n_cap - num_nodes):
            n_cap += self.node_step
        return n_cap, e_cap

    @staticmethod
    def for_batches(batches: Sequence[Batch], node_step: int = 256, edge_step: int = 512) -> "BucketPolicy":
        """Caps measured over a set of collated batches (a stand-in for dataset statistics)."""
        mn = max(int(b["max_nodes_per_graph"]) for b in batches)
        me = 0
        for b in batches:
            if "max_edges_per_graph" in b:
                me = max(me, int(b["max_edges_per_graph"]))
            else:
                me = max(me, int(torch.bincount(b.batch[b.edge_index[0]], minlength=int(b.num_graphs)).max()))
        return BucketPolicy(max(mn, 2), max(me, 2), node_step, edge_step)


@dataclass(frozen=True)
class _Shape:
    """Everything a captured graph bakes in."""
    n_cap: int
    e_cap: int
    graphs: int                 # B real graphs
    dummies: int                # D trailing dummy graphs
    max_nodes: int
    max_edges: int              # 0: no block-diagonal promise (general radix CSR path)
    no_self_loops: bool


class StagedBatch:
    """One mini-batch packed for upload: x | edge_index | batch | y in ONE pinned host buffer (256-byte aligned
    sub-ranges, padded to its bucket), so the per-step upload is a single host-to-device copy."""

    def __init__(self, shape: _Shape, buf: Tensor, views: Dict[str, Tensor], num_nodes: int, num_edges: int,
                 raw: bool = False, upload_bytes: Optional[int] = None):
        self.shape, self.buf, self.views = shape, buf, views
        self.num_nodes, self.num_edges = num_nodes, num_edges        # real (unpadded) sizes
        self.raw = raw                # True: edge_index holds per-graph LOCAL indices, `batch` is not filled --
        #                               the device derives both from `counts` (ghscn_collate_batch)
        self.upload_bytes = int(buf.numel()) if upload_bytes is None else int(upload_bytes)
        self.device_copy: Optional[Tensor] = None                    # set by GraphHSCNStep.make_resident

    @property
    def nbytes(self) -> int:
        """Bytes that travel host -> device for this batch."""
        return self.upload_bytes


def _layout(shape: _Shape, x_like: Tensor, y_like: Tensor) -> Tuple[Dict[str, tuple], int]:
    """name -> (byte offset, dtype, shape) of the packed buffer; total bytes.  `batch` comes last: a batch that is
    collated on the device does not upload it (everything in front of it is the upload)."""
    bt = shape.graphs + shape.dummies
    specs = {"x": (x_like.dtype, (shape.n_cap,) + tuple(x_like.shape[1:])),
             "edge_index": (torch.int64, (2, shape.e_cap)),
             "y": (y_like.dtype, (bt,) + tuple(y_like.shape[1:])),
             "counts": (torch.int32, (2, bt)),                    # nodes per graph, edges per graph
             "batch": (torch.int64, (shape.n_cap,))}
    out, total = {}, 0
    for k, (dt, shp) in specs.items():
        nbytes = int(torch.empty((), dtype=dt).element_size()) * int(math.prod(shp))
        out[k] = (total, dt, shp)
        total += (nbytes + 255) // 256 * 256
    return out, max(total, 256)


def _views(buf: Tensor, layout: Dict[str, tuple]) -> Dict[str, Tensor]:
    out = {}
    for k, (off, dt, shp) in layout.items():
        nbytes = int(torch.empty((), dtype=dt).element_size()) * int(math.prod(shp))
        out[k] = buf[off:off + nbytes].view(dt).view(shp)
    return out


def _dummy_plan(shape: _Shape, n_real: int, e_real: int) -> Tuple[List[int], List[int]]:
    """(nodes, edges) of each of the D trailing dummy graphs: pad nodes fill the dummy graphs in order (each within the
    per-graph cap), pad edges go to the first dummy graphs that have at least two nodes."""
    D = shape.dummies
    pn, pe = shape.n_cap - n_real, shape.e_cap - e_real
    sizes, left = [], pn
    for _ in range(D):
        take = min(left, shape.max_nodes)
        sizes.append(take)
        left -= take
    if left:
        raise ValueError(f"{pn} pad nodes do not fit {D} dummy graphs of <= {shape.max_nodes} nodes")
    edges = []
    for sz in sizes:
        take = 0
        if pe > 0 and sz >= 2:
            take = min(pe, shape.max_edges) if shape.max_edges else pe
        edges.append(take)
        pe -= take
    if pe > 0:
        raise ValueError(f"pad edges do not fit the dummy graphs (left {pe})")
    return sizes, edges


def _fill_dummies(views: Dict[str, Tensor], shape: _Shape, n_real: int, e_real: int, local: bool = False) -> None:
    """Writes the D trailing dummy graphs: zero features / labels, pad edges as (a, a+1), (a+1, a) pairs walking along
    each dummy graph's nodes (global node ids, or ids local to the dummy graph with `local`)."""
    B, D = shape.graphs, shape.dummies
    sizes, edges = _dummy_plan(shape, n_real, e_real)
    views["x"][n_real:].zero_()
    views["y"][B:].zero_()
    if not local:
        views["batch"][n_real:] = torch.repeat_interleave(torch.arange(B, B + D), torch.tensor(sizes))
    views["counts"][0, B:] = torch.tensor(sizes, dtype=torch.int32)
    views["counts"][1, B:] = torch.tensor(edges, dtype=torch.int32)
    ei = views["edge_index"]
    pos, base = e_real, n_real
    for sz, take in zip(sizes, edges):
        if take:
            k = torch.arange(take)
            a = (0 if local else base) + (k // 2) % (sz - 1)
            fwd = (k % 2 == 0)
            ei[0, pos:pos + take] = torch.where(fwd, a, a + 1)
            ei[1, pos:pos + take] = torch.where(fwd, a + 1, a)
        pos, base = pos + take, base + sz


def shape_for(batch: Batch, policy: Optional[BucketPolicy], use_blocks: bool = True) -> _Shape:
    """The static shape a collated batch runs under: its bucket (with a policy) or its exact sizes (without)."""
    N, E, B = int(batch.x.size(0)), int(batch.edge_index.size(1)), int(batch.num_graphs)
    loop_free = not bool((batch.edge_index[0] == batch.edge_index[1]).any())
    max_n = int(batch["max_nodes_per_graph"]) if "max_nodes_per_graph" in batch else int(
        (batch.ptr[1:] - batch.ptr[:-1]).max())
    max_e = 0
    if use_blocks:
        if "max_edges_per_graph" in batch:                  # verified by Batch.from_data_list at collate time
            max_e = int(batch["max_edges_per_graph"])
        else:
            blocks = edge_blocks_from_batch(batch.edge_index, batch.batch, B)
            max_e = blocks[1] if blocks is not None else 0
    if policy is None:
        return _Shape(N, E, B, 0, max_n, max_e, loop_free)
    if max_n > policy.max_nodes_per_graph or max_e > policy.max_edges_per_graph:
        raise ValueError(f"batch has a graph with {max_n} nodes / {max_e} edges, above the policy's caps "
                         f"({policy.max_nodes_per_graph} / {policy.max_edges_per_graph})")
    n_cap, e_cap = policy.bucket(N, E)
    return _Shape(n_cap, e_cap, B, policy.dummy_graphs, policy.max_nodes_per_graph,
                  policy.max_edges_per_graph if max_e else 0, loop_free)


def stage_batch(batch: Batch, policy: Optional[BucketPolicy], pin: bool = False, use_blocks: bool = True
                ) -> StagedBatch:
    """Packs one collated batch (x, edge_index, batch, y) into a single host buffer of its bucket's size; the padding
    is written as trailing dummy graphs (`_fill_dummies`)."""
    shape = shape_for(batch, policy, use_blocks)
    layout, nbytes = _layout(shape, batch.x, batch.y)
    buf = torch.empty(nbytes, dtype=torch.uint8)
    if pin:
        buf = buf.pin_memory()
    v = _views(buf, layout)
    N, E = int(batch.x.size(0)), int(batch.edge_index.size(1))
    v["x"][:N].copy_(batch.x)
    v["edge_index"][:, :E].copy_(batch.edge_index)
    v["batch"][:N].copy_(batch.batch)
    v["y"][:shape.graphs].copy_(batch.y)
    v["counts"][0, :shape.graphs] = (batch.ptr[1:] - batch.ptr[:-1]).to(torch.int32)
    v["counts"][1, :shape.graphs] = torch.bincount(batch.batch[batch.edge_index[0]], minlength=shape.graphs).to(torch.int32)
    if shape.dummies:
        _fill_dummies(v, shape, N, E)
    return StagedBatch(shape, buf, v, N, E)


def stage_graphs(graphs: Sequence, policy: Optional[BucketPolicy], pin: bool = False) -> StagedBatch:
    """Collate WITHOUT the host-side index arithmetic (loader/loader.py:48-60): the graphs' x, LOCAL edge_index and y
    are written back to back straight into one (pinned) staging buffer together with the per-graph node / edge
    counts; `batch` and the offset edge_index are derived on the device (ghscn_collate_batch) after the upload.
    Usable as a DataLoader `collate_fn` through `StagingCollate`."""
    B = len(graphs)
    nodes = torch.tensor([int(g.num_nodes) for g in graphs], dtype=torch.int64)
    edges = torch.tensor([int(g.edge_index.size(1)) for g in graphs], dtype=torch.int64)
    N, E = int(nodes.sum()), int(edges.sum())
    max_n, max_e = int(nodes.max()), int(edges.max())
    ei_local = torch.cat([g.edge_index for g in graphs], dim=1) if E else torch.zeros((2, 0), dtype=torch.int64)
    if E:
        bound = torch.repeat_interleave(nodes, edges)
        if int(ei_local.min()) < 0 or bool((ei_local >= bound).any()):
            raise ValueError("a graph has an edge that leaves it (edge_index out of range)")
    loop_free = not bool((ei_local[0] == ei_local[1]).any())
    if policy is None:
        shape = _Shape(N, E, B, 0, max_n, max_e, loop_free)
    else:
        if max_n > policy.max_nodes_per_graph or max_e > policy.max_edges_per_graph:
            raise ValueError(f"a graph has {max_n} nodes / {max_e} edges, above the policy's caps "
                             f"({policy.max_nodes_per_graph} / {policy.max_edges_per_graph})")
        n_cap, e_cap = policy.bucket(N, E)
        shape = _Shape(n_cap, e_cap, B, policy.dummy_graphs, policy.max_nodes_per_graph,
                       policy.max_edges_per_graph if max_e else 0, loop_free)
    x0, y0 = graphs[0].x, graphs[0].y
    layout, nbytes = _layout(shape, x0, y0)
    buf = torch.empty(nbytes, dtype=torch.uint8)
    if pin:
        buf = buf.pin_memory()
    v = _views(buf, layout)
    torch.cat([g.x for g in graphs], dim=0, out=v["x"][:N])
    v["edge_index"][:, :E].copy_(ei_local)
    torch.cat([g.y for g in graphs], dim=0, out=v["y"][:B])
    v["counts"][0, :B] = nodes.to(torch.int32)
    v["counts"][1, :B] = edges.to(torch.int32)
    if shape.dummies:
        _fill_dummies(v, shape, N, E, local=True)
    return StagedBatch(shape, buf, v, N, E, raw=True, upload_bytes=layout["batch"][0])


class StagingCollate:
    """`collate_fn` for torch's DataLoader: a list of graphs -> a StagedBatch in pinned memory, ready for
    `GraphHSCNStep.load` (device-side collate).  The reference's counterpart is PyG's Python collate behind
    `DataLoader(dataset, batch_size=...)` at loader/loader.py:48-60."""

    def __init__(self, policy: Optional[BucketPolicy], pin: bool = True):
        self.policy, self.pin = policy, pin

    def __call__(self, graphs) -> StagedBatch:
        return stage_graphs(list(graphs), self.policy, pin=self.pin)


# ---------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------
TWO_STREAMS = os.environ.get("GHSCN_TWO_STREAMS", "1") != "0"
# data parallel: exchange the gradients of everything behind the first layer while that layer's backward still runs
OVERLAP_ALLREDUCE = os.environ.get("GHSCN_OVERLAP_ALLREDUCE", "1") != "0"


@dataclass
class StepConfig:
    num_features: int = 9
    num_classes: int = 10
    num_clusters: int = 10
    scn_units: Tuple[int, ...] = (16,)
    scn_act: str = "elu"
    hidden: int = 300
    num_layers: int = 3
    activation: str = "relu"
    loss_fn: str = "cross_entropy"
    lr: float = 1e-3
    weight_decay: float = 5e-4
    clip_grad_norm: bool = False        # train/train.py:92-93 (max_norm 1.0)
    batch_accumulation: int = 1         # train/train.py:89-91: optimizer step every this many batches (HSCN stage)


class _Runner:
    """Static device buffers + captured CUDA graphs of one shape bucket."""

    def __init__(self, shape: _Shape, layout: Dict[str, tuple], nbytes: int, device: torch.device):
        self.shape = shape
        if device.type == "cuda":
            self.dev_buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self.dev = _views(self.dev_buf, layout)
        else:
            self.dev_buf = None
            self.dev = {k: torch.empty(shp, dtype=dt) for k, (off, dt, shp) in layout.items()}
        self.hints = dict(num_graphs=shape.graphs + shape.dummies, batch_sorted=1, max_nodes_per_graph=shape.max_nodes,
                          no_self_loops=int(shape.no_self_loops))
        self.graphs: Dict[tuple, "torch.cuda.CUDAGraph"] = {}
        self._raw: Optional[Dict[str, Tensor]] = None

    @property
    def raw_dev(self) -> Dict[str, Tensor]:
        """Tensors of a batch that is collated on the device: x / y / counts and the LOCAL edge list live in the
        uploaded buffer, `edge_index` (offset) and `batch` are written by ghscn_collate_batch into their own buffers."""
        if self._raw is None:
            d = self.dev
            self._raw = dict(x=d["x"], y=d["y"], counts=d["counts"], edge_index_local=d["edge_index"],
                             edge_index=torch.empty_like(d["edge_index"]), batch=torch.empty_like(d["batch"]))
        return self._raw


class GraphHSCNStep:
    """Owns the two models, their optimizers, and per-bucket static device buffers + CUDA graphs.

        step = GraphHSCNStep(cfg, first_batch, device, policy=BucketPolicy(...))
        for batch in loader:                       # variable-shape batches
            step.load(batch)                       # pack (pinned) + upload + select the bucket
            step.run()                             # replays the bucket's graph (captured on first use)
            losses = step.download()

    Without a policy the single "bucket" has the first batch's exact shape (no padding, D = 0)."""

    def __init__(self, cfg: StepConfig, host_batch: Batch, device: torch.device, op_ns: Optional[SimpleNamespace] = None,
                 seed: int = 0, padded: bool = True, policy: Optional[BucketPolicy] = None, auto_capture: bool = False):
        from . import pyg
        self.cfg, self.device, self.padded, self.policy = cfg, device, padded, policy
        self.ns = op_ns or pyg.namespace()
        self.B = int(host_batch.num_graphs)
        self.auto_capture = auto_capture and device.type == "cuda" and padded
        self.world = 1
        self._use_blocks = os.environ.get("GHSCN_BLOCKED_CSR", "1") != "0"
        self._runners: Dict[_Shape, _Runner] = {}
        self._pool = None
        self._x_like, self._y_like = host_batch.x[:0], host_batch.y[:0]
        self._copy_stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self._slots: List[Optional[Tensor]] = [None, None]      # device staging buffers of the prefetch pipeline
        self._slot_events: List[Optional["torch.cuda.Event"]] = [None, None]
        self._next_slot = 0
        self._micro = 0                                         # batches since the last HSCN optimizer step
        self._warmed: set = set()                               # step variants that have run eagerly at least once
        self._loss_ring: Optional[List[Tensor]] = None
        self._loss_slot = 0
        self._prep_stream = None
        torch.manual_seed(seed)
        self.scn = models.SCN(list(cfg.scn_units), cfg.scn_act, cfg.num_features, cfg.num_clusters, ops=self.ns).to(device)
        self.hscn = models.HSCN("GAT", "GCN", "GCN", models.ACTIVATIONS[cfg.activation], cfg.num_features, cfg.hidden,
                                cfg.num_classes, cfg.num_layers, ops=self.ns).to(device)
        self.losses = torch.zeros(3, dtype=torch.float32, device=device)     # mincut, ortho, task
        self.losses_host = torch.zeros(3, dtype=torch.float32).pin_memory() if device.type == "cuda" else torch.zeros(3)
        self.staged = self.stage(host_batch)
        self.runner = self._runner_for(self.staged.shape)
        self.upload()
        self._prepare()

    # -- compatibility views of the current bucket -------------------------------------------------------------
    @property
    def dev(self) -> Dict[str, Tensor]:
        return self.runner.raw_dev if self.staged.raw else self.runner.dev

    @property
    def host(self) -> Dict[str, Tensor]:
        return self.staged.views

    @property
    def hints(self) -> Dict[str, int]:
        return self.runner.hints

    @property
    def h2d_bytes(self) -> int:
        return self.staged.nbytes

    @property
    def graph(self):
        return self.runner.graphs.get(self._variant())

    @property
    def num_buckets(self) -> int:
        return len(self._runners)

    @property
    def num_graphs_captured(self) -> int:
        return sum(len(r.graphs) for r in self._runners.values())

    # -- host side: pack a batch into its bucket -------------------------------------------------------------------
    def stage(self, batch: Batch) -> StagedBatch:
        """Host-side packing of one collated batch into a pinned buffer of its bucket's size (padding = dummy graphs)."""
        if int(batch.num_graphs) != self.B:
            raise ValueError(f"this step was built for {self.B} graphs per batch, got {int(batch.num_graphs)}")
        return stage_batch(batch, self.policy, pin=self.device.type == "cuda", use_blocks=self._use_blocks)

    def stage_graphs(self, graphs: Sequence) -> StagedBatch:
        """Packs a list of graphs without host-side collate arithmetic; `batch` / offset edge_index are derived on the
        device at the start of the step (see `stage_graphs`)."""
        if len(graphs) != self.B:
            raise ValueError(f"this step was built for {self.B} graphs per batch, got {len(graphs)}")
        return stage_graphs(graphs, self.policy, pin=self.device.type == "cuda")

    def _runner_for(self, shape: _Shape) -> _Runner:
        r = self._runners.get(shape)
        if r is None:
            layout, nbytes = _layout(shape, self._x_like, self._y_like)
            r = self._runners[shape] = _Runner(shape, layout, nbytes, self.device)
        return r

    # -- host <-> device ---------------------------------------------------------------------------
    def load(self, batch) -> StagedBatch:
        """Packs (unless already a StagedBatch), selects the bucket and uploads on the current stream."""
        staged = batch if isinstance(batch, StagedBatch) else self.stage(batch)
        self.select(staged)
        self.upload()
        return staged

    def select(self, staged: StagedBatch) -> None:
        self.staged = staged
        self.runner = self._runner_for(staged.shape)

    def upload(self) -> None:
        """Host -> device copy of the current staged batch into its bucket's static buffer (current stream)."""
        r = self.runner
        if r.dev_buf is not None:
            n = self.staged.nbytes                  # a device-collated batch does not upload `batch`
            r.dev_buf[:n].copy_(self.staged.buf[:n], non_blocking=True)
            return
        for k, v in self.staged.views.items():
            r.dev[k].copy_(v)

    def make_resident(self, staged: StagedBatch) -> StagedBatch:
        """Keeps a device copy of the packed batch (benchmarks with inputs already in HBM)."""
        staged.device_copy = staged.buf[:staged.nbytes].to(self.device)
        return staged

    def select_resident(self, staged: StagedBatch) -> None:
        """Device -> device copy of a resident batch into its bucket's static buffer."""
        self.select(staged)
        self.runner.dev_buf[:staged.nbytes].copy_(staged.device_copy, non_blocking=True)

    def prefetch(self, staged: StagedBatch) -> int:
        """Starts the H2D copy of a LATER step's batch on the copy stream (double-buffered device staging slots): it
        overlaps the step that is running.  -> slot to hand to `select_prefetched`."""
        slot = self._next_slot
        self._next_slot ^= 1
        cur = torch.cuda.current_stream()
        buf = self._slots[slot]
        if buf is None or buf.numel() < staged.nbytes:
            buf = self._slots[slot] = torch.empty(max(staged.nbytes, 1 << 22), dtype=torch.uint8, device=self.device)
        self._copy_stream.wait_stream(cur)          # the slot's previous consumer (a D2D on `cur`) must be done
        with torch.cuda.stream(self._copy_stream):
            buf[:staged.nbytes].copy_(staged.buf[:staged.nbytes], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._slot_events[slot] = ev
        return slot

    def select_prefetched(self, staged: StagedBatch, slot: int) -> None:
        self.select(staged)
        torch.cuda.current_stream().wait_event(self._slot_events[slot])
        self.runner.dev_buf[:staged.nbytes].copy_(self._slots[slot][:staged.nbytes], non_blocking=True)

    def download(self) -> Tensor:
        self.losses_host.copy_(self.losses, non_blocking=True)
        return self.losses_host

    def download_async(self) -> Tuple[Tensor, Optional["torch.cuda.Event"]]:
        """D2H read of the three losses into one of two alternating pinned buffers; -> (host tensor, event to wait
        for).  Lets a training loop read step t's losses while step t+1 is already running."""
        if self.device.type != "cuda":
            return self.download(), None
        if self._loss_ring is None:
            self._loss_ring = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_slot ^= 1
        host = self._loss_ring[self._loss_slot]
        host.copy_(self.losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev

    # -- one-time preparation: materialise lazy parameters, find live parameters, build optimizers ------
    def _forward_scn(self, x_f: Tensor):
        d = self.dev
        ei, ew = self.ns.gcn_norm(d["edge_index"], None, x_f.size(0), add_self_loops=True)
        return (ei, ew) + tuple(self._scn_losses(x_f, ei, ew))

    def _scn_losses(self, x_f, ei, ew, losses_tensor: bool = False):
        # with dummy graphs the MinCUT losses are taken over the B real graphs only
        real = self.B if self.runner.shape.dummies else None
        return self.scn.forward_batched(x_f, ei, ew, self.dev["batch"], losses_tensor=losses_tensor,
                                        num_graphs=real)

    def _assign(self, x_f: Tensor, ei: Tensor, ew: Tensor):
        d, shape = self.dev, self.runner.shape
        with torch.no_grad():
            s = self.scn.logits(x_f, ei, ew)
            clusters = hetero.assign_clusters(torch.softmax(s, dim=-1))
        return hetero.build_hetero_batch(d["x"], d["edge_index"], d["batch"], clusters,
                                         self.cfg.num_clusters, y=d["y"], padded=self.padded,
                                         num_graphs=shape.graphs + shape.dummies, x_float=x_f)

    def _hscn_loss(self, hb) -> Tensor:
        """pred = HSCN(batch); loss = criterion(pred, y) (train/train.py:81-82).  On CUDA, for [B, C] float targets and
        <= 256 graphs, the output layer, the loss and their backward are one kernel (the step differentiates the loss
        itself, so the incoming gradient is 1)."""
        y = hb["local"].y
        lin = self.hscn.lin_2
        if (self.device.type == "cuda" and y.dim() == 2 and self.cfg.loss_fn in ops.LOSS_MODES
                and os.environ.get("GHSCN_FUSED_HEAD", "1") != "0" and torch.is_floating_point(y)
                and ops.head_out_loss_ok(self.runner.shape.graphs + self.runner.shape.dummies, lin.weight.size(1),
                                         lin.weight.size(0))):
            hidden = self.hscn.forward_hidden(hb.x_dict, hb.edge_index_dict, hb)
            return ops.head_out_loss(self.cfg.loss_fn, hidden, lin.weight, lin.bias, y, rows=self.B,
                                     unit_grad=True)[0]
        pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)
        return self._task_loss(pred, y)

    def _task_loss(self, pred: Tensor, y: Tensor) -> Tensor:
        """criterion(loss_fn, pred, true) of train/train.py:82 over the B real graphs (the dummy graphs' rows carry no
        loss and no gradient); on CUDA the loss, its gradient and the sigmoid score come from one kernel."""
        if (pred.is_cuda and pred.dim() == 2 and y.dim() == 2 and self.cfg.loss_fn in ops.LOSS_MODES
                and os.environ.get("GHSCN_FUSED_LOSS", "1") != "0"):
            return ops.graph_loss(self.cfg.loss_fn, pred, y, rows=self.B)[0]
        if self.runner.shape.dummies:
            pred, y = pred[:self.B], y[:self.B]
        return models.criterion(self.cfg.loss_fn, pred, y)[0]

    def _prepare(self) -> None:
        cfg = self.cfg
        with structure_hints(**self.hints):
            self._collate_on_device()
            self._register_blocks()
            x_f = self._cast(self.dev["x"])
            ei, ew, _, mc, ol = self._forward_scn(x_f)
            scn_live = live_parameter_names(self.scn, mc + ol)
            hb = self._assign(x_f, ei, ew)
            loss = self._hscn_loss(hb)                                     # materialises lazy weights
            hscn_live = live_parameter_names(self.hscn, loss)
        self.scn_grads = FlatGradients(self.scn, scn_live)
        self.hscn_grads = FlatGradients(self.hscn, hscn_live)
        kw = dict(lr=cfg.lr, weight_decay=cfg.weight_decay)
        if self.device.type == "cuda":
            self.scn_opt = FlatAdamW(self.scn_grads.flatten_parameters().detach(), self.scn_grads.flat, **kw)
            self.hscn_opt = FlatAdamW(self.hscn_grads.flatten_parameters().detach(), self.hscn_grads.flat, **kw)
        else:
            self.scn_opt = torch.optim.AdamW([self.scn_grads.flatten_parameters()], **kw)
            self.hscn_opt = torch.optim.AdamW([self.hscn_grads.flatten_parameters()], **kw)
        structure_cache().clear()

    def _register_blocks(self) -> None:
        shape = self.runner.shape
        if shape.max_edges and self.device.type == "cuda":
            d = self.dev
            seg = structure_cache().segments(d["batch"], shape.graphs + shape.dummies)
            structure_cache().register_blocks(d["edge_index"], seg.ptr, shape.graphs + shape.dummies, shape.max_nodes,
                                              shape.max_edges)

    def _cast(self, x: Tensor) -> Tensor:
        if x.dtype == torch.int64 and x.is_cuda:
            return ops.cast_i64_f32(x)
        return x.float()

    # -- optimizer-side pieces shared by both schedules ---------------------------------------------------
    def _variant(self) -> tuple:
        """(accumulate onto the HSCN gradient buffer, take the HSCN optimizer step) for the batch about to run."""
        acc = max(int(self.cfg.batch_accumulation), 1)
        return (self._micro > 0, self._micro + 1 >= acc, bool(self.staged.raw))

    def _hscn_backward(self, loss: Tensor, world: int, accumulate: bool, update: bool) -> bool:
        """Backward of the HSCN stage into the flat gradient buffer; -> True if the data-parallel exchange is already
        done (overlapped with the first layer's backward, NCCL only)."""
        b = getattr(self.hscn, "first_layer_output", None)
        self.hscn.first_layer_output = None
        if (update and not accumulate and world > 1 and OVERLAP_ALLREDUCE and b is not None and b.requires_grad
                and b.is_cuda and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"):
            split = self.hscn_grads.leading("convs.0.")
            if 0 < split < len(self.hscn_grads.params):
                self.hscn_grads.backward_reduce_overlapped(loss, b, split)
                return True
        self.hscn_grads.backward_into(loss, accumulate=accumulate)
        return False

    def _hscn_update(self, world: int, reduced: bool = False) -> None:
        if not reduced:
            self.hscn_grads.all_reduce_mean(world)
        if self.cfg.clip_grad_norm:
            if isinstance(self.hscn_opt, FlatAdamW):
                self.hscn_opt.clip_grad_norm(1.0)
            else:
                nn.utils.clip_grad_norm_(self.hscn_grads.params, 1.0)
        self.hscn_opt.step()

    # -- the three stages ---------------------------------------------------------------------------------
    def _step_serial(self, world: int, accumulate: bool, update: bool) -> None:
        self._register_blocks()
        x_f = self._cast(self.dev["x"])
        ei, ew, _, mc, ol = self._forward_scn(x_f)
        self.scn_grads.backward_into(mc + ol)
        self.losses[0:1].copy_(mc.detach().view(1))
        self.losses[1:2].copy_(ol.detach().view(1))
        self.scn_grads.all_reduce_mean(world)
        self.scn_opt.step()
        hb = self._assign(x_f, ei, ew)
        loss = self._hscn_loss(hb)
        reduced = self._hscn_backward(loss, world, accumulate, update)
        self.losses[2:3].copy_(loss.detach().view(1))
        if update:
            self._hscn_update(world, reduced)
        if self._prep_stream is not None:
            torch.cuda.current_stream().wait_stream(self._prep_stream)

    def _step_two_streams(self, world: int, accumulate: bool, update: bool) -> None:
        """Same operations as `_step_serial`, scheduled on two CUDA streams.  The "local" half of the HSCN (l->l convs,
        readout, loss, backward, AdamW) depends on neither the SCN stage nor the cluster assignment: only the
        "virtual" branch does (model/hscn.py:84-94 has no virtual->local relation).  So the SCN step, the assignment
        (K7) and the virtual branch run on the branch stream while the caller's stream runs the local half; the two
        meet once, at the end of the step.  Everything both halves read is produced on the caller's stream BEFORE
        the fork (float features, `ptr`, the plain CSR of the molecular graph); every kernel and its inputs are the
        same as in the serial schedule, so results are bit-identical (tests/test_gpu_step.py).  The deferred join is
        valid because nothing on the caller's stream reads a virtual-branch tensor; tensors that cross streams are
        marked with `record_stream` (pyg/nn.py), so the allocator cannot recycle them under a reader."""
        from .pyg import nn as pnn
        main = torch.cuda.current_stream()
        side = pnn.branch_stream(self.device)
        pnn.take_forked_streams(self.device)                     # forget forks of earlier (already joined) work
        d, shape = self.dev, self.runner.shape
        N = d["x"].size(0)
        x_f = self._cast(d["x"])
        structure_cache().segments(d["batch"], shape.graphs + shape.dummies)
        self._register_blocks()
        plain = structure_cache().graph(d["edge_index"], N, N, False)
        plain.by_dst, plain.by_src                               # built here, read by both streams
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ei, ew = self.ns.gcn_norm(d["edge_index"], None, N, add_self_loops=True)
            _, both = self._scn_losses(x_f, ei, ew, losses_tensor=True)
            self.scn_grads.backward_into(both.sum())     # == (mincut + ortho).backward(), one reduction instead of
            self.losses[0:2].copy_(both.detach())        # two select/scatter round trips
            self.scn_grads.all_reduce_mean(world)
            self.scn_opt.step()
            hb = self._assign(x_f, ei, ew)
        self.hscn.defer_branch_join = True
        try:
            loss = self._hscn_loss(hb)
        finally:
            self.hscn.defer_branch_join = False
        reduced = self._hscn_backward(loss, world, accumulate, update)
        self.losses[2:3].copy_(loss.detach().view(1))
        if update:
            self._hscn_update(world, reduced)
        main.wait_stream(side)
        for st in pnn.take_forked_streams(self.device):          # every stream the HeteroConv layers forked
            if st is not side:
                main.wait_stream(st)
        if self._prep_stream is not None:
            main.wait_stream(self._prep_stream)

    def _prefetch_weight_images(self) -> None:
        """tcgen05 weight images of this step, built on a side stream right away (they depend on the parameters only)."""
        from . import gemm
        gemm.clear_images()
        # off by default: measured +8 us per step (the image kernels then compete with the first kernels of the local
        # chain for SMs; in front of their GEMM they sit in that chain's idle launch gaps)
        if self.device.type != "cuda" or os.environ.get("GHSCN_PREFETCH_IMAGES", "0") == "0":
            return
        if self._prep_stream is None:
            self._prep_stream = torch.cuda.Stream(device=self.device)
        self._prep_stream.wait_stream(torch.cuda.current_stream())
        gemm.prefetch_images(self._prep_stream)

    def _step(self, world: int, variant: Optional[tuple] = None) -> None:
        accumulate, update = (variant if variant is not None else self._variant())[:2]
        self._prefetch_weight_images()
        self._collate_on_device()
        if TWO_STREAMS and self.device.type == "cuda":
            self._step_two_streams(world, accumulate, update)
        else:
            self._step_serial(world, accumulate, update)

    def _collate_on_device(self) -> None:
        """`batch` and the offset edge_index of a raw-staged batch (loader/loader.py:48-60 on the device)."""
        if not self.staged.raw:
            return
        from ._lib import lib
        from .structure import _p, _stream
        d, shape = self.runner.raw_dev, self.runner.shape
        g = shape.graphs + shape.dummies
        lib().call("ghscn_collate_batch", _p(d["counts"][0]), _p(d["counts"][1]), g, _p(d["edge_index_local"]),
                   shape.e_cap, _p(d["edge_index"]), _p(d["batch"]), shape.n_cap, _stream())

    def _advance(self) -> None:
        self._micro = 0 if self._variant()[1] else self._micro + 1

    def predict(self) -> Tensor:
        """HSCN logits [B, C] of the current batch with the current weights (no gradient, no update): the forward of
        train/train.py:115-123 `eval_epoch`.  Eager; the cluster assignment comes from the current SCN weights."""
        with torch.no_grad(), structure_hints(**self.hints):
            structure_cache().clear()
            self._collate_on_device()
            d = self.dev
            self._register_blocks()
            x_f = self._cast(d["x"])
            ei, ew = self.ns.gcn_norm(d["edge_index"], None, x_f.size(0), add_self_loops=True)
            hb = self._assign(x_f, ei, ew)
            pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)
        return pred[:self.B]

    def _snapshot(self):
        opts = [o for o in (self.scn_opt, self.hscn_opt) if isinstance(o, FlatAdamW)]
        return [t.clone() for o in opts for t in (o.p, o.g, o.exp_avg, o.exp_avg_sq, o._state)] + [self.losses.clone()]

    def _restore(self, snap) -> None:
        opts = [o for o in (self.scn_opt, self.hscn_opt) if isinstance(o, FlatAdamW)]
        dst = [t for o in opts for t in (o.p, o.g, o.exp_avg, o.exp_avg_sq, o._state)] + [self.losses]
        for t, v in zip(dst, snap):
            t.copy_(v)

    def _dry_run(self, world: int, variant: tuple) -> None:
        """One eager step whose effects are rolled back (parameters, gradients, optimizer state, losses): loads every
        kernel / library handle the step uses, so a capture that follows never meets a lazy initialisation."""
        snap = self._snapshot()
        with structure_hints(**self.hints):
            structure_cache().clear()
            self._step(world, variant)
        self._restore(snap)
        torch.cuda.synchronize()
        self._warmed.add(variant)

    def run_eager(self, world: int = 1) -> None:
        variant = self._variant()
        with structure_hints(**self.hints):
            structure_cache().clear()
            self._step(world, variant)
        self._warmed.add(variant)
        self._advance()

    # -- CUDA graph capture: static shapes (padded virtual layout), no host sync inside -------------------
    def capture(self, world: int = 1, warmup: int = 3, variant: Optional[tuple] = None) -> None:
        """Captures the current bucket's graph for `variant` (default: the variant of the batch about to run).
        `warmup` eager steps run first (they are real training steps on the current batch)."""
        assert self.device.type == "cuda" and self.padded, "capture needs CUDA and the padded virtual layout"
        self.world = world
        variant = variant or self._variant()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                with structure_hints(**self.hints):
                    structure_cache().clear()
                    self._step(world, variant)
            if warmup:
                self._warmed.add(variant)
            elif variant not in self._warmed:
                self._dry_run(world, variant)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        structure_cache().check_blocked_status(self.device)     # the per-graph CSR promise held on real data
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with capture_scope(), structure_hints(**self.hints):
            with torch.cuda.graph(g, pool=self._pool):
                self._step(world, variant)
        self.runner.graphs[variant] = g

    def run(self, world: int = 1) -> None:
        variant = self._variant()
        g = self.runner.graphs.get(variant)
        if g is None and self.auto_capture:
            # first batch of this bucket: capture without warm-up steps (warm-up would train on the batch more than
            # once; the very first capture does one rolled-back dry run instead)
            self.capture(world or self.world, warmup=0, variant=variant)
            g = self.runner.graphs[variant]
        if g is not None:
            g.replay()
            self._advance()
        else:
            self.run_eager(world)

    def release_graphs(self) -> None:
        for r in self._runners.values():
            r.graphs.clear()
