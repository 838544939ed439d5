"""Times the h=300 GCN aggregation kernel alone (CUDA graph of 40 launches, rotating operands > L2) on several inputs:
   python scripts/spmm_probe.py"""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graph_hscn_b200 import synthetic
from graph_hscn_b200._lib import lib
from graph_hscn_b200.data import Batch
from graph_hscn_b200.structure import _p, _stream, structure_cache, structure_hints
from graph_hscn_b200.train import BucketPolicy, stage_batch

dev = torch.device("cuda")


def time_spmm(ei, N, hidden=300, reps=40, nset=10, label=""):
    structure_cache().clear()
    st = structure_cache().graph(ei, N, N, False)
    w, w_t, _ = st.weights(None, normalize=True)
    d = st.by_dst
    xs = [torch.randn(N, hidden, device=dev) for _ in range(nset)]
    ys = [torch.empty(N, hidden, device=dev) for _ in range(nset)]
    L = lib()

    def launch(i, s):
        L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xs[i % nset]), hidden, _p(ys[i % nset]), hidden, None,
               N, hidden, 0, s)
    for i in range(nset):
        launch(i, _stream())
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = _stream()
        for i in range(reps):
            launch(i, s)
    g.replay(); torch.cuda.synchronize()
    tr = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tr.append(a.elapsed_time(b) / reps * 1e3)
    nnz = d.num_items
    by = 4 * hidden * 2 * N + 8 * nnz + 4 * (N + 1)
    us = statistics.mean(tr)
    deg = (d.rowptr[1:] - d.rowptr[:-1])
    print(f"{label:40s} N={N} nnz={nnz} maxdeg={int(deg.max())} {us:7.2f} us  {by / us / 1e3:7.1f} GB/s  frac {by / us / 1e3 / 6544.7:.3f}",
          flush=True)


b_old = synthetic.peptides_batch(128, seed=1236)
time_spmm(b_old.edge_index.to(dev), b_old.x.size(0), label="round-1 batch (sequential seed 1236)")
b_new = Batch.from_data_list([synthetic.peptides_graph(1236 + 6, i) for i in range(128)])
time_spmm(b_new.edge_index.to(dev), b_new.x.size(0), label="indexed batch 6, exact")
st = stage_batch(b_new, BucketPolicy(444, 1024))
time_spmm(st.views["edge_index"].to(dev), st.shape.n_cap, label="indexed batch 6, bucketed (dummy graph)")
time_spmm(b_old.edge_index.to(dev), b_old.x.size(0), label="round-1 batch again")
