"""Graph-HSCN training step on one B200 (optionally CUDA-graph captured) and its data-parallel form.

One step = one pass of the hot path over one `batch`/`ptr` mini-batch (BASELINE config #2):
    1. SCN stage   -- gcn_norm(add_self_loops) -> GraphConv stack -> logits -> fused MinCUT losses,
                      backward, AdamW step                    (train/train_clustering.py:34-50, batched)
    2. assignment  -- SCN forward with the updated weights, softmax, first-max cluster id
                      (train_clustering.py:57-69) and on-device virtual-node construction
                      (loader/hetero_data.py:42-87)
    3. HSCN stage  -- 3-relation HeteroConv stack, mean readout, 2 linears, loss, backward, AdamW step
                      (train/train.py:73-95)
Data parallelism: graphs are independent, so ranks take disjoint contiguous graph
ranges and exchange nothing in forward/backward; one NCCL all-reduce of a single flat fp32 gradient
buffer per model per step averages the gradients (the reference has no distributed code).
"""
from __future__ import annotations

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
    balance the summed weight instead of the graph count (SURVEY 8e "balance by sum n_g")."""
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

    def backward_into(self, loss: Tensor) -> None:
        """`loss.backward()` for the live parameters, with the gradients written straight into the flat buffer: the
        autograd engine hands the finished gradients over (`torch.autograd.grad`) and ONE multi-tensor copy places them,
        instead of one `grad += g` kernel per parameter interleaved with the backward chain (and no zero fill before).
        Every value is the one `backward()` would have accumulated onto a zeroed buffer."""
        grads = torch.autograd.grad(loss, self.params)
        torch._foreach_copy_([p.grad for p in self.params], list(grads))

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

    def all_reduce_mean(self, world: Optional[int] = None, group=None) -> None:
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = world or dist.get_world_size(group)
        if world == 1:
            return
        if os.environ.get("GHSCN_SKIP_ALLREDUCE") == "1":      # timing experiments only: WRONG gradients
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.flat.mul_(1.0 / world)


class FlatAdamW:
    """torch.optim.AdamW semantics on one flat parameter buffer, one kernel per step (ghscn_adamw_step); the step
    counter lives on the device, so `step()` is CUDA-graph capturable."""

    def __init__(self, flat_param: Tensor, flat_grad: Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2):
        self.p, self.g = flat_param, flat_grad
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat_grad)
        self.exp_avg_sq = torch.zeros_like(flat_grad)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=flat_grad.device)
        self.state = {0: {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_count}}

    def step(self) -> None:
        from ._lib import lib
        from .structure import _p, _stream
        lib().call("ghscn_adamw_step", _p(self.p), _p(self.g), _p(self.exp_avg), _p(self.exp_avg_sq), self.p.numel(),
                   float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd),
                   _p(self.step_count), _stream())

    def zero_grad(self) -> None:
        self.g.zero_()


def live_parameter_names(module: nn.Module, loss: Tensor) -> List[str]:
    """Names of the parameters reachable from `loss` (consistent on all ranks for the same model)."""
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, [p for _, p in named], allow_unused=True, retain_graph=False)
    return [n for (n, _), g in zip(named, grads) if g is not None]


# ---------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------
TWO_STREAMS = os.environ.get("GHSCN_TWO_STREAMS", "1") != "0"


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


class GraphHSCNStep:
    """Owns the two models, their optimizers and static device buffers for one batch shape."""

    def __init__(self, cfg: StepConfig, host_batch: Batch, device: torch.device, op_ns: Optional[SimpleNamespace] = None,
                 seed: int = 0, padded: bool = True):
        from . import pyg
        self.cfg, self.device, self.padded = cfg, device, padded
        self.ns = op_ns or pyg.namespace()
        self.B = int(host_batch.num_graphs)
        counts = host_batch.ptr[1:] - host_batch.ptr[:-1]
        self.hints = dict(num_graphs=self.B, batch_sorted=1, max_nodes_per_graph=int(counts.max()),
                          no_self_loops=int(not bool((host_batch.edge_index[0] == host_batch.edge_index[1]).any())))
        # collate-time fact: edges are graph-major and block diagonal -> per-graph CSR kernel (K1 fast path)
        blocks = edge_blocks_from_batch(host_batch.edge_index, host_batch.batch, self.B)
        self.max_edges_per_graph = blocks[1] if blocks is not None and os.environ.get("GHSCN_BLOCKED_CSR", "1") != "0" else 0
        # pinned host staging + static device buffers (inputs are re-copied every step in the e2e path)
        src = {k: host_batch[k].contiguous() for k in ("x", "edge_index", "batch", "y")}
        if device.type == "cuda":
            # ONE pinned staging buffer and ONE device buffer hold all four inputs (256-byte aligned sub-ranges), so
            # the per-step upload is a single host-to-device copy; `host[k]` / `dev[k]` are typed views into them
            offs, total = {}, 0
            for k, v in src.items():
                offs[k] = total
                total += (v.numel() * v.element_size() + 255) // 256 * 256
            self._host_buf = torch.empty(max(total, 256), dtype=torch.uint8).pin_memory()
            self._dev_buf = torch.empty(max(total, 256), dtype=torch.uint8, device=device)

            def view(buf, k, v):
                nbytes = v.numel() * v.element_size()
                return buf[offs[k]:offs[k] + nbytes].view(v.dtype).view(v.shape)
            self.host = {k: view(self._host_buf, k, v) for k, v in src.items()}
            for k, v in src.items():
                self.host[k].copy_(v)
            self.dev = {k: view(self._dev_buf, k, v) for k, v in src.items()}
            self.h2d_bytes = int(self._host_buf.numel())
        else:
            self._host_buf = self._dev_buf = None
            self.host = src
            self.dev = {k: torch.empty_like(v, device=device) for k, v in self.host.items()}
            self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        torch.manual_seed(seed)
        self.scn = models.SCN(list(cfg.scn_units), cfg.scn_act, cfg.num_features, cfg.num_clusters, ops=self.ns).to(device)
        self.hscn = models.HSCN("GAT", "GCN", "GCN", models.ACTIVATIONS[cfg.activation], cfg.num_features, cfg.hidden,
                                cfg.num_classes, cfg.num_layers, ops=self.ns).to(device)
        self.losses = torch.zeros(3, dtype=torch.float32, device=device)     # mincut, ortho, task
        self.losses_host = torch.zeros(3, dtype=torch.float32).pin_memory() if device.type == "cuda" else torch.zeros(3)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.upload()
        self._prepare()

    # -- host <-> device ---------------------------------------------------------------------------
    def upload(self) -> None:
        if self._dev_buf is not None:
            self._dev_buf.copy_(self._host_buf, non_blocking=True)
            return
        for k, v in self.host.items():
            self.dev[k].copy_(v, non_blocking=True)

    def download(self) -> Tensor:
        self.losses_host.copy_(self.losses, non_blocking=True)
        return self.losses_host

    # -- one-time preparation: materialise lazy parameters, find live parameters, build optimizers ------
    def _forward_scn(self, x_f: Tensor):
        ei, ew = self.ns.gcn_norm(self.dev["edge_index"], None, x_f.size(0), add_self_loops=True)
        return (ei, ew) + tuple(self.scn.forward_batched(x_f, ei, ew, self.dev["batch"]))

    def _assign(self, x_f: Tensor, ei: Tensor, ew: Tensor):
        with torch.no_grad():
            s = self.scn.logits(x_f, ei, ew)
            clusters = hetero.assign_clusters(torch.softmax(s, dim=-1))
        return hetero.build_hetero_batch(self.dev["x"], self.dev["edge_index"], self.dev["batch"], clusters,
                                         self.cfg.num_clusters, y=self.dev["y"], padded=self.padded,
                                         num_graphs=self.B, x_float=x_f)

    def _prepare(self) -> None:
        cfg = self.cfg
        with structure_hints(**self.hints):
            x_f = self._cast(self.dev["x"])
            ei, ew, _, mc, ol = self._forward_scn(x_f)
            scn_live = live_parameter_names(self.scn, mc + ol)
            hb = self._assign(x_f, ei, ew)
            pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)            # materialises lazy weights
            loss, _ = models.criterion(cfg.loss_fn, pred, hb["local"].y)
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
        if self.max_edges_per_graph and self.device.type == "cuda":
            seg = structure_cache().segments(self.dev["batch"], self.B)
            structure_cache().register_blocks(self.dev["edge_index"], seg.ptr, self.B,
                                              self.hints["max_nodes_per_graph"], self.max_edges_per_graph)

    def _cast(self, x: Tensor) -> Tensor:
        if x.dtype == torch.int64 and x.is_cuda:
            return ops.cast_i64_f32(x)
        return x.float()

    # -- the three stages; `sync_grads` hooks the data-parallel all-reduce in between ------------------
    def stage_scn_backward(self):
        self.scn_grads.zero()
        self._register_blocks()
        x_f = self._cast(self.dev["x"])
        ei, ew, _, mc, ol = self._forward_scn(x_f)
        (mc + ol).backward()
        self.losses[0:1].copy_(mc.detach().view(1))
        self.losses[1:2].copy_(ol.detach().view(1))
        return x_f, ei, ew

    def stage_assign_and_hscn_backward(self, x_f, ei, ew):
        self.scn_opt.step()
        hb = self._assign(x_f, ei, ew)
        self.hscn_grads.zero()
        pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)
        loss, _ = models.criterion(self.cfg.loss_fn, pred, hb["local"].y)
        loss.backward()
        self.losses[2:3].copy_(loss.detach().view(1))

    def stage_hscn_update(self):
        self.hscn_opt.step()

    def _step_serial(self, world: int) -> None:
        st = self.stage_scn_backward()
        self.scn_grads.all_reduce_mean(world)
        self.stage_assign_and_hscn_backward(*st)
        self.hscn_grads.all_reduce_mean(world)
        self.stage_hscn_update()

    def _step_two_streams(self, world: int) -> None:
        """Same operations as `_step_serial`, scheduled on two CUDA streams.  The "local" half of the HSCN (l->l convs,
        readout, loss, backward, AdamW) depends on neither the SCN stage nor the cluster assignment: only the
        "virtual" branch does (model/hscn.py:84-94 has no virtual->local relation).  So the SCN step, the assignment
        (K7) and the virtual branch run on the branch stream while the caller's stream runs the local half; the two
        meet once, at the end of the step.  Everything both halves read is produced on the caller's stream BEFORE
        the fork (float features, `ptr`, the plain CSR of the molecular graph); every kernel and its inputs are the
        same as in the serial schedule, so results are bit-identical (tests/test_gpu_step.py)."""
        from .pyg import nn as pnn
        main = torch.cuda.current_stream()
        side = pnn.branch_stream(self.device)
        N = self.dev["x"].size(0)
        x_f = self._cast(self.dev["x"])
        structure_cache().segments(self.dev["batch"], self.B)
        self._register_blocks()
        plain = structure_cache().graph(self.dev["edge_index"], N, N, False)
        plain.by_dst, plain.by_src                               # built here, read by both streams
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ei, ew = self.ns.gcn_norm(self.dev["edge_index"], None, N, add_self_loops=True)
            _, both = self.scn.forward_batched(x_f, ei, ew, self.dev["batch"], losses_tensor=True)
            self.scn_grads.backward_into(both.sum())     # == (mincut + ortho).backward(), one reduction instead of
            self.losses[0:2].copy_(both.detach())        # two select/scatter round trips
            self.scn_grads.all_reduce_mean(world)
            self.scn_opt.step()
            hb = self._assign(x_f, ei, ew)
        self.hscn.defer_branch_join = True
        try:
            pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)
        finally:
            self.hscn.defer_branch_join = False
        loss, _ = models.criterion(self.cfg.loss_fn, pred, hb["local"].y)
        self.hscn_grads.backward_into(loss)
        self.losses[2:3].copy_(loss.detach().view(1))
        self.hscn_grads.all_reduce_mean(world)
        self.hscn_opt.step()
        main.wait_stream(side)

    def _step(self, world: int) -> None:
        if TWO_STREAMS and self.device.type == "cuda":
            self._step_two_streams(world)
        else:
            self._step_serial(world)

    def run_eager(self, world: int = 1) -> None:
        with structure_hints(**self.hints):
            structure_cache().clear()
            self._step(world)

    # -- CUDA graph capture: static shapes (padded virtual layout), no host sync inside -------------------
    def capture(self, world: int = 1, warmup: int = 3) -> None:
        assert self.device.type == "cuda" and self.padded, "capture needs CUDA and the padded virtual layout"
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.run_eager(world)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with capture_scope(), structure_hints(**self.hints):
            with torch.cuda.graph(self.graph):
                self._step(world)

    def run(self, world: int = 1) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager(world)
