"""CPU restatement of one Graph-HSCN step (the benchmark's `cpu_baseline` / `--impl reference` arm).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the reference's own control flow for one
mini-batch of graphs:
    train/train_clustering.py:34-50   SCN step: gcn_norm -> SCN -> (mc_loss + o_loss).backward() -> optimizer
                                      (batched over the mini-batch; `per_graph=True` reproduces the
                                      reference's one-optimizer-step-per-graph loop exactly)
    train/train_clustering.py:57-69   assignment pass: softmax(s).max(1)[1].cpu().numpy() per graph
    loader/hetero_data.py:42-87       per-graph Python/numpy HeteroData construction
    loader/hetero_data.py:91-104      collate
    train/train.py:73-95              HSCN forward, criterion, backward, optimizer step
All arithmetic is the pure-torch oracle (oracle/ops.py, oracle/nn.py) on CPU threads.
"""
from __future__ import annotations

from typing import List

import torch

from graph_hscn_b200 import models
from graph_hscn_b200.data import Batch, Data, HeteroData

from . import hetero as ohet
from .namespace import namespace


class OracleStep:
    def __init__(self, cfg, host_batch: Batch, seed: int = 0, per_graph: bool = False):
        self.cfg, self.per_graph = cfg, per_graph
        self.ns = namespace()
        self.batch = host_batch
        self.graphs: List[Data] = host_batch.to_data_list()
        torch.manual_seed(seed)
        self.scn = models.SCN(list(cfg.scn_units), cfg.scn_act, cfg.num_features, cfg.num_clusters, ops=self.ns)
        self.hscn = models.HSCN("GAT", "GCN", "GCN", models.ACTIVATIONS[cfg.activation], cfg.num_features,
                                cfg.hidden, cfg.num_classes, cfg.num_layers, ops=self.ns)
        hb = self._hetero(self._clusters())
        self.hscn(hb.x_dict, hb.edge_index_dict, hb)          # materialise lazy parameters
        kw = dict(lr=cfg.lr, weight_decay=cfg.weight_decay)
        self.scn_opt = torch.optim.AdamW(self.scn.parameters(), **kw)
        self.hscn_opt = torch.optim.AdamW(self.hscn.parameters(), **kw)
        self.losses = [0.0, 0.0, 0.0]
        self._micro = 0
        self.hscn_opt.zero_grad()

    def _scn_train(self) -> None:
        o, b = self.ns, self.batch
        if self.per_graph:
            for d in self.graphs:
                ei, ew = o.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
                self.scn_opt.zero_grad()
                _, mc, ol, _ = self.scn(d.x.float(), ei, ew)
                (mc + ol).backward()
                self.scn_opt.step()
        else:
            ei, ew = o.gcn_norm(b.edge_index, None, b.x.size(0), add_self_loops=True)
            self.scn_opt.zero_grad()
            _, mc, ol = self.scn.forward_batched(b.x.float(), ei, ew, b.batch)
            (mc + ol).backward()
            self.scn_opt.step()
        self.losses[0], self.losses[1] = float(mc.detach()), float(ol.detach())
        self.scn_grads = {n: p.grad.detach().clone() for n, p in self.scn.named_parameters() if p.grad is not None}

    def _clusters(self) -> list:
        out = []
        for d in self.graphs:           # train_clustering.py:57-69 (no torch.no_grad in the reference)
            ei, ew = self.ns.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
            clust, _, _, _ = self.scn(d.x.float(), ei, ew)
            out.append(ohet.cluster_argmax(clust))
        return out

    def _hetero(self, clusters: list):
        hl = []
        for d, c in zip(self.graphs, clusters):
            _, vx, vv, lv = ohet.virtual_nodes(d.x, c, self.cfg.num_clusters)
            h = HeteroData()
            h["local"].x = d.x.float()
            h["local"].y = d.y
            h["virtual"].x = vx
            h["local", "to", "local"].edge_index = d.edge_index
            h["virtual", "to", "virtual"].edge_index = vv
            h["local", "to", "virtual"].edge_index = lv
            hl.append(h)
        return Batch.from_data_list(hl)

    def run(self) -> None:
        self._scn_train()
        hb = self._hetero(self._clusters())
        pred = self.hscn(hb.x_dict, hb.edge_index_dict, hb)
        loss, _ = models.criterion(self.cfg.loss_fn, pred, hb["local"].y)
        loss.backward()
        self.pred = pred.detach()
        self.grads = {n: p.grad.detach().clone() for n, p in self.hscn.named_parameters() if p.grad is not None}
        self._micro += 1
        if self._micro % max(int(getattr(self.cfg, "batch_accumulation", 1)), 1) == 0:     # train/train.py:89-95
            if getattr(self.cfg, "clip_grad_norm", False):
                torch.nn.utils.clip_grad_norm_(self.hscn.parameters(), 1.0)
            self.hscn_opt.step()
            self.hscn_opt.zero_grad()
        self.losses[2] = float(loss.detach())

    def set_batch(self, host_batch: Batch) -> None:
        """Next mini-batch of the loader (variable shapes are free on the CPU path)."""
        self.batch = host_batch
        self.graphs = host_batch.to_data_list()
