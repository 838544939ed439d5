"""BASELINE.json configs #1, #3, #4, #5 on one B200: kernel sweeps and model steps (markdown to stdout / file).

    python scripts/sweeps.py [out.md]

Kernel times are device times: N back-to-back launches captured in one CUDA graph, operands rotated over more
memory than the 126 MB L2 where the working set allows.  GB/s = algorithmic bytes / time, against
the measured copy peak of MEASURED_PEAKS.json.
"""
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_hscn_b200.ops  # noqa: F401,E402
from graph_hscn_b200 import models, pyg, synthetic  # noqa: E402
from graph_hscn_b200._lib import lib  # noqa: E402
from graph_hscn_b200.data import Batch  # noqa: E402
from graph_hscn_b200.structure import _p, _stream, structure_cache, structure_hints  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out_lines = []


def emit(s=""):
    print(s, flush=True)
    out_lines.append(s)


def graph_time(fn, reps=20, trials=5):
    """us per call of fn(i), `reps` calls captured in one CUDA graph."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(trials):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return statistics.median(ts)


# ------------------------------------------------------------------------------------------------ SpMM sweep
def spmm_sweep():
    emit("## Config #5a: SpMM (K2) avg-degree sweep, N = 153,600 nodes, directed ER edges\n")
    emit("| avg degree | F | nnz | us / launch | algorithmic MB | GB/s | frac of measured HBM peak |")
    emit("|---|---|---|---|---|---|---|")
    N = 153600
    for deg in (2, 8, 32):
        g = torch.Generator().manual_seed(5 + deg)
        e = int(deg * N)
        ei = torch.stack([torch.randint(0, N, (e,), generator=g), torch.randint(0, N, (e,), generator=g)]).to(dev)
        st = structure_cache().graph(ei, N, N, False)
        w, _, _ = st.weights(None, normalize=True)
        d = st.by_dst
        for F in (64, 300, 512):
            nset = max(2, int(600e6 // (2 * 4 * F * N)) + 1)
            xs = [torch.randn(N, F, device=dev) for _ in range(nset)]
            ys = [torch.empty(N, F, device=dev) for _ in range(nset)]
            L = lib()

            def fn(i):
                L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xs[i % nset]), F, _p(ys[i % nset]), F, None,
                       N, F, 0, _stream())
            us = graph_time(fn)
            by = 4 * F * 2 * N + 8 * e + 4 * (N + 1)
            emit(f"| {deg} | {F} | {e} | {us:.1f} | {by / 1e6:.1f} | {by / us / 1e3:.0f} | {by / us / 1e3 / PEAK:.2f} |")
            del xs, ys
        structure_cache().clear()
    emit()


# ---------------------------------------------------------------------------------------------- MinCUT sweep
def mincut_sweep():
    emit("## Config #5b: fused MinCUT pool (K6), B = 1024 Peptides-shaped graphs (one CTA per graph)\n")
    emit("| K | H | fwd us | fwd+bwd us | algorithmic MB (fwd) | fwd GB/s | frac | fwd GFLOP/s |")
    emit("|---|---|---|---|---|---|---|---|")
    b = synthetic.peptides_batch(1024, seed=1239)
    N = b.x.size(0)
    o_ei = b.edge_index.to(dev)
    batch = b.batch.to(dev)
    counts = (b.ptr[1:] - b.ptr[:-1])
    hints = dict(num_graphs=1024, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)
    with structure_hints(**hints):
        ei, _ = pyg.gcn_norm(o_ei, None, N, add_self_loops=True)
        E1 = ei.size(1)
        for K in (4, 16, 64, 128):
            for H in (64, 300, 512):
                g = torch.Generator().manual_seed(K * H)
                x = torch.randn(N, H, generator=g).to(dev)
                s = torch.randn(N, K, generator=g).to(dev).requires_grad_()

                def fwd(i):
                    return pyg.mincut_pool_ragged(x, ei, s, batch)
                try:
                    us_f = graph_time(lambda i: fwd(i), reps=5)

                    def fb(i):
                        out, adj, mc, ol = fwd(i)
                        (mc + ol + out.sum() * 1e-3 + adj.sum()).backward()
                        s.grad = None
                    us_fb = graph_time(fb, reps=3)
                except Exception as ex:       # shapes outside the shared-memory / workspace design
                    emit(f"| {K} | {H} | unsupported: {str(ex)[:60]} | | | | | |")
                    continue
                by = 4 * N * K * 2 + 4 * N * H + 4 * (N + 1) + 4 * E1 + 1024 * (4 * K * H + 8 * K * K + 32)
                fl = 2 * N * K * H + 2 * E1 * K + 4 * N * K * K + 2 * N * K
                emit(f"| {K} | {H} | {us_f:.0f} | {us_fb:.0f} | {by / 1e6:.1f} | {by / us_f / 1e3:.0f} | "
                     f"{by / us_f / 1e3 / PEAK:.2f} | {fl / us_f / 1e3:.0f} |")
    structure_cache().clear()
    emit()


# ------------------------------------------------------------------------------------------ VOC-SP (config #4)
def vocsp():
    emit("## Config #4: PascalVOC-SP-shaped graphs (n in [395,500], avg degree ~5.7, 14 feats), K=32, H=256, operator level\n")
    emit("| B | nodes | GCNConv 14->256 + MinCUT K=32 fwd us | fwd+bwd us | graphs/s (fwd+bwd) |")
    emit("|---|---|---|---|---|")
    for B in (32, 128):
        graphs = synthetic.vocsp_graphs(B, seed=1238)
        b = Batch.from_data_list(graphs)
        N = b.x.size(0)
        counts = (b.ptr[1:] - b.ptr[:-1])
        x, ei, batch = b.x.to(dev), b.edge_index.to(dev), b.batch.to(dev)
        torch.manual_seed(0)
        conv = pyg.GCNConv(14, 256).to(dev)
        lin = pyg.Linear(256, 32).to(dev)
        hints = dict(num_graphs=B, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)
        with structure_hints(**hints):
            def fwd(i):
                h = torch.relu(conv(x, ei))
                return pyg.mincut_pool_ragged(h, ei, lin(h), batch)
            us_f = graph_time(lambda i: fwd(i), reps=5)

            def fb(i):
                out, adj, mc, ol = fwd(i)
                (mc + ol + out.sum() * 1e-3 + adj.sum()).backward()
                for p in list(conv.parameters()) + list(lin.parameters()):
                    p.grad = None
            us_fb = graph_time(fb, reps=3)
        emit(f"| {B} | {N} | {us_f:.0f} | {us_fb:.0f} | {B / us_fb * 1e6:.0f} |")
        structure_cache().clear()
    emit()


# ---------------------------------------------------------------------------------------- model steps #1, #3
def model_steps():
    emit("## Config #1 (MPNN GCN L=5 h=300, B=128) and config #3 shape (Graph-HSCN step, B=1024, 11 targets, L1)\n")
    emit("| config | graphs/step | ms/step (CUDA graph replay) | graphs/s |")
    emit("|---|---|---|---|")
    # config #1
    b = synthetic.peptides_batch(128, seed=1235)
    b.x = b.x.float()
    bd = b.to(dev)
    torch.manual_seed(0)
    m = models.MPNN("gcn", torch.relu, 9, 300, 10, 5, dropout=0.0).to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=5e-4, fused=True, capturable=True)
    counts = (b.ptr[1:] - b.ptr[:-1])
    hints = dict(num_graphs=128, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)

    def step(i):
        structure_cache().clear()
        bd.register_structures()                 # collate-time facts: ptr, block-diagonal edge list (K1 fast path)
        opt.zero_grad(set_to_none=False)
        loss, _ = models.criterion("cross_entropy", m(bd), bd.y)
        loss.backward()
        opt.step()
    with structure_hints(**hints):
        for i in range(3):
            step(i)      # grads exist before capture
        us = graph_time(step, reps=5)
    emit(f"| #1 MPNN(GCN) L=5 h=300 fwd+bwd+AdamW | 128 | {us / 1e3:.3f} | {128 / us * 1e6:.0f} |")
    structure_cache().clear()
    # config #3 shape on one GPU
    from graph_hscn_b200.train import GraphHSCNStep, StepConfig
    hb = synthetic.peptides_batch(1024, seed=1237, task="struct")
    st = GraphHSCNStep(StepConfig(num_classes=11, loss_fn="l1"), hb, dev, padded=True)
    st.capture(warmup=2)
    for _ in range(3):
        st.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.run(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ms = statistics.median(ts)
    emit(f"| #3 Graph-HSCN step, Peptides-struct shape (N={hb.x.size(0)}) | 1024 | {ms:.3f} | {1024 / ms * 1e3:.0f} |")
    emit()


if __name__ == "__main__":
    emit(f"# Round-1 sweeps on one B200 (measured HBM copy peak {PEAK:.0f} GB/s)\n")
    which = os.environ.get("SWEEPS", "spmm,mincut,vocsp,models").split(",")
    t0 = time.time()
    if "spmm" in which:
        spmm_sweep()
    if "mincut" in which:
        mincut_sweep()
    if "vocsp" in which:
        vocsp()
    if "models" in which:
        model_steps()
    emit(f"_total wall time {time.time() - t0:.0f} s_")
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write("\n".join(out_lines) + "\n")
