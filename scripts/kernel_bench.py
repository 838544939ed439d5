"""Stand-alone device times of the small / fused kernels at the bench shape (each timed alone: `reps` launches in one
CUDA graph, no other stream active).   python scripts/kernel_bench.py"""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import graph_hscn_b200.ops as ops  # noqa
from graph_hscn_b200 import gemm, hetero, synthetic
from graph_hscn_b200._lib import lib
from graph_hscn_b200.data import Batch
from graph_hscn_b200.structure import _p, _stream, structure_cache

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
L = lib()


def graph_time(fn, reps=20, trials=5):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(trials):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return statistics.median(ts)


def report(name, us, note=""):
    print(f"{name:58s} {us:8.2f} us  {note}", flush=True)


g = torch.Generator().manual_seed(0)
# ---- small linear
for (m, k, n, act) in [(129, 300, 300, 2), (129, 300, 10, 0), (1290, 300, 300, 2), (1290, 9, 300, 0), (1027, 300, 300, 2)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
    y = torch.empty(m, n, device=dev); dy = torch.randn(m, n, device=dev); dx = torch.empty(m, k, device=dev)
    dw = torch.empty(n, k, device=dev); db = torch.empty(n, device=dev)
    wsb = L.query("ghscn_small_linear_dw_workspace_bytes", m, k, n)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
    report(f"small_linear fwd  [{m}x{k}] -> {n}", graph_time(lambda i: L.call("ghscn_small_linear_fwd", _p(x), k, _p(w), k, _p(b), act, m, k, n, _p(y), n, _stream())))
    report(f"small_linear dx   [{m}x{k}] -> {n}", graph_time(lambda i: L.call("ghscn_small_linear_dx", _p(dy), n, _p(y), n, act, _p(w), k, m, k, n, _p(dx), k, _stream())))
    report(f"small_linear dw   [{m}x{k}] -> {n}", graph_time(lambda i: L.call("ghscn_small_linear_dw", _p(dy), n, _p(y), n, act, _p(x), k, m, k, n, _p(dw), _p(db), _p(ws), wsb, _stream())))
    report(f"torch F.linear    [{m}x{k}] -> {n}", graph_time(lambda i: torch.nn.functional.linear(x, w, b)))
m, k, n = 1290, 300, 300
x1 = torch.randn(m, k, device=dev); x2 = torch.randn(m, k, device=dev)
w1 = torch.randn(n, k, device=dev); w2 = torch.randn(n, k, device=dev); b1 = torch.randn(n, device=dev); y = torch.empty(m, n, device=dev)
report("small_linear2 fwd [1290x(300+300)] -> 300 relu", graph_time(lambda i: L.call("ghscn_small_linear2_fwd", _p(x1), k, _p(w1), k, _p(b1), k, _p(x2), k, _p(w2), k, _p(b1), k, 2, m, n, _p(y), n, _stream())))

# ---- fused attention pool at the bench shape
b = Batch.from_data_list([synthetic.peptides_graph(1236, i) for i in range(128)]).to(dev)
N = b.x.size(0)
for mode in ("random clusters", "one big cluster per graph"):
    if mode == "random clusters":
        clusters = torch.randint(0, 10, (N,), generator=g).int().to(dev)
    else:
        clusters = (torch.rand(N, generator=g) < 0.9).int().to(dev) * 3
    hb = hetero.build_hetero_batch(b.x, b.edge_index, b.batch, clusters, 10, padded=True, num_graphs=128)
    lv = hb["local", "to", "virtual"].edge_index
    V = hb["virtual"].x.size(0)
    st = structure_cache().graph(lv, N, V, False)
    d = st.by_dst
    for F in (9, 300):
        xs = torch.randn(N, F, device=dev); xd = torch.randn(V, F, device=dev)
        u = torch.randn(2, F, device=dev); pooled = torch.empty(V, F, device=dev)
        wsrc = torch.randn(300, F, device=dev); att = torch.randn(300, device=dev)
        for wpr in (1, 2, 4, 8):
            report(f"gat_pool_fused F={F} warps/row={wpr} ({mode})", graph_time(lambda i: L.call("ghscn_gat_pool_fused_fwd", _p(d.rowptr), _p(d.col), _p(xs), F, _p(xd), F, _p(u[0]), _p(u[1]), 0.2, V, F, 1, _p(pooled), F, wpr, _stream())),
                   f"maxlen {int((d.rowptr[1:] - d.rowptr[:-1]).max())}")
        report(f"gat_fold_attention F={F}", graph_time(lambda i: L.call("ghscn_gat_fold_attention", _p(wsrc), F, _p(att), _p(wsrc), F, _p(att), 300, F, F, 1, _p(u[0]), _p(u[1]), _stream())))
# ---- loss
pred = torch.randn(129, 10, device=dev); yt = torch.rand(129, 10, device=dev)
loss = torch.empty(1, device=dev); dp = torch.empty(129, 10, device=dev); sc = torch.empty(129, 10, device=dev)
report("graph_loss 129x10", graph_time(lambda i: L.call("ghscn_graph_loss", _p(pred), 10, _p(yt), 10, 128, 129, 10, 0, _p(loss), _p(dp), _p(sc), _stream())))
