"""MinCUT pool at the bench shape (B = 128 Peptides graphs, K = 10): forward and forward+backward device time, alone on
the GPU.  PROBE_H=0 (default): losses only (what hscn.py:63 keeps), on 16 features; PROBE_H=300: with the pooled
features and coarse adjacency and their gradients.  PROBE_GRAPHS / PROBE_K pick the batch and cluster count;
GHSCN_MINCUT_THREADS=256|512|1024 selects the CTA size, GHSCN_MINCUT_STREAM_X=0 / GHSCN_MINCUT_STAGE=1 the A/B
switches of csrc/mincut.cu.  python scripts/mincut_small_probe.py"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import pyg, synthetic  # noqa: E402
from graph_hscn_b200.structure import structure_cache, structure_hints  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("PROBE_GRAPHS", "128"))
K = int(os.environ.get("PROBE_K", "10"))
H = int(os.environ.get("PROBE_H", "0"))
b = synthetic.peptides_batch(B, seed=1239)
N = b.x.size(0)
counts = b.ptr[1:] - b.ptr[:-1]
hints = dict(num_graphs=B, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)


def graph_time(fn, reps=20, trials=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(trials):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3 / reps)
    return statistics.median(ts)


with structure_hints(**hints):
    ei, _ = pyg.gcn_norm(b.edge_index.to(dev), None, N, add_self_loops=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H or 16, generator=g).to(dev).requires_grad_(H > 0)
    s = torch.randn(N, K, generator=g).to(dev).requires_grad_()
    batch = b.batch.to(dev)

    def fwd():
        return pyg.mincut_pool_ragged(x, ei, s, batch, want_out=H > 0, want_adj=H > 0)

    def fb():
        out, adj, mc, ol = fwd()
        loss = mc + ol
        if H > 0:
            loss = loss + out.sum() + adj.sum()
        loss.backward()
        s.grad = None
        x.grad = None
    tf = graph_time(fwd) if os.environ.get('PROBE_SKIP_FWD') != '1' else float('nan')
    tfb = graph_time(fb)
print(f"B={B} K={K} H={H} N={N} stream_x={os.environ.get('GHSCN_MINCUT_STREAM_X', '1')} "
      f"stage={os.environ.get('GHSCN_MINCUT_STAGE', 'default')} threads={os.environ.get('GHSCN_MINCUT_THREADS', 'auto')}: fwd {tf:.1f} us, fwd+bwd {tfb:.1f} us")
structure_cache().clear()
