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
    # correctness: against index_add_ in fp64 (weights w are per slot, sources col, destination = row of the slot)
    dst = torch.repeat_interleave(torch.arange(N, device=dev), (d.rowptr[1:] - d.rowptr[:-1]).long())
    ref = torch.zeros(N, hidden, device=dev, dtype=torch.float64)
    nz = int(d.rowptr[-1])
    ref.index_add_(0, dst, xs[0][d.col[:nz].long()].double() * w[:nz].double()[:, None])
    err = float((ys[0].double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-6, f"SpMM result wrong: {err}"
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
