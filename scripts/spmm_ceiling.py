"""Where is the ceiling for a 2 x 21.9 MB streaming kernel on this GPU? (copy vs identity-SpMM vs real SpMM)"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graph_hscn_b200._lib import lib
from graph_hscn_b200.structure import _p, _stream, structure_cache, build_csr
import graph_hscn_b200.ops
dev = torch.device("cuda:0")
b = bench.make_batch(0)
N, F = b.x.size(0), 300
ei = b.edge_index.to(dev)
st = structure_cache().graph(ei, N, N, False)
w, _, _ = st.weights(None, normalize=True)
d = st.by_dst
ident = torch.arange(N, device=dev)
idc = build_csr(ident, ident, N, False)
wi = torch.ones(N, device=dev)
nset = 10
xs = [torch.randn(N, F, device=dev) for _ in range(nset)]
ys = [torch.empty(N, F, device=dev) for _ in range(nset)]
L, s_ = lib(), _stream()
def t(fn, reps=40):
    for i in range(nset): fn(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, c) in enumerate(ev):
        a.record(); fn(i); c.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(c) for a, c in ev) * 1e3
def t_batch(fn, reps=40):   # one event pair around `reps` back-to-back launches: amortises launch/event overhead
    for i in range(nset): fn(i)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) * 1e3 / reps
spmm = lambda i: L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xs[i % nset]), F, _p(ys[i % nset]), F, None, N, F, 0, s_)
spmm_id = lambda i: L.call("ghscn_spmm", _p(idc.rowptr), _p(idc.col), _p(wi), _p(xs[i % nset]), F, _p(ys[i % nset]), F, None, N, F, 0, s_)
copy = lambda i: ys[i % nset].copy_(xs[i % nset])
empty = lambda i: L.call("ghscn_cast_i64_f32", _p(ident), 1, _p(wi), s_)
for name, fn in [("empty kernel", empty), ("torch copy 21.9MB->21.9MB", copy), ("identity spmm (1 slot/row)", spmm_id), ("gcn spmm (2.0 slots/row)", spmm)]:
    print(f"{name:32s} per-launch events {t(fn):7.2f} us   back-to-back {t_batch(fn):7.2f} us")

def t_graph(fn, reps=40):
    """`reps` launches captured in one CUDA graph: device-side back-to-back, no CPU launch overhead."""
    for i in range(nset): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(g):
        global s_
        s_ = _stream()
        for i in range(reps): fn(i)
    s_ = _stream()
    g.replay(); torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) * 1e3 / reps
print("--- CUDA-graph replay (device back-to-back) ---")
for name, fn in [("empty kernel", empty), ("torch copy 21.9MB->21.9MB", copy), ("identity spmm (1 slot/row)", spmm_id), ("gcn spmm (2.0 slots/row)", spmm)]:
    print(f"{name:32s} graph replay {t_graph(fn):7.2f} us/launch")
