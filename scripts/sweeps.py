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
    emit("## Config #5b: MinCUT pool (K6), B = 1024 Peptides-shaped graphs (K <= 16: fused per-graph kernel + batch-wide "
         "streaming of the pooled features; K >= 32: split phases with the contractions on the tensor cores)\n")
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


# ----------------------------------------------------------------------- per-kernel roofline at the bench shape
def kernel_table():
    """Every kernel family of SURVEY 8a at the config #2 shape (B = 128 Peptides graphs, h = 300, K = 10), timed alone:
    20 back-to-back calls in one CUDA graph, so launch gaps are excluded but nothing else overlaps."""
    from graph_hscn_b200 import gemm, hetero, ops
    from graph_hscn_b200.structure import StructureCache, build_csr, edge_blocks_from_batch
    emit("## Per-kernel roofline at the bench shape (config #2: B = 128, h = 300, K = 10), each kernel timed alone\n")
    b = synthetic.peptides_batch(128, seed=1236)
    N, E, B, H, K, F = b.x.size(0), b.edge_index.size(1), 128, 300, 10, 9
    emit(f"N = {N} nodes, E = {E} directed edges, V = B*K = {B * K} virtual slots; peak = measured HBM copy {PEAK:.0f} GB/s\n")
    emit("| kernel (SURVEY 8a row) | us | algorithmic MB | GB/s | frac of HBM peak | note |")
    emit("|---|---|---|---|---|---|")
    ei, batch, x_raw = b.edge_index.to(dev), b.batch.to(dev), b.x.to(dev)
    counts = b.ptr[1:] - b.ptr[:-1]
    hints = dict(num_graphs=B, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)
    blocks = edge_blocks_from_batch(b.edge_index, b.batch, B)

    def row(name, us, by, note=""):
        emit(f"| {name} | {us:.1f} | {by / 1e6:.2f} | {by / us / 1e3:.0f} | {by / us / 1e3 / PEAK:.3f} | {note} |")

    with structure_hints(**hints):
        cache = StructureCache()
        seg = cache.segments(batch, B)

        def k1_fast(i):
            c = StructureCache()
            c.register_blocks(ei, seg.ptr, B, *blocks)
            c.graph(ei, N, N, False).by_dst
        row("K1 CSR, per-graph fast path, both orientations (a1)", graph_time(k1_fast), 16 * E + 8 * (N + 1) + 16 * E,
            "one launch; latency-bound (128 CTAs, 7 dependent phases)")

        def k1_radix(i):
            build_csr(ei[1], ei[0], N, False)
            build_csr(ei[0], ei[1], N, False)
        row("K1 CSR, general radix path, both orientations (a1)", graph_time(k1_radix), 2 * (16 * E + 4 * (N + 1) + 8 * E),
            "16 launches")
        st = structure_cache().graph(ei, N, N, True)
        d, t = st.by_dst, st.by_src

        def k1b(i):
            st._weights.clear()
            st.weights(None, normalize=True, need_transpose=False)
        row("K1b gcn_norm on the CSR: deg^-1/2 + slot weights (a1)", graph_time(k1b), 12 * d.num_items + 8 * N, "2 launches")
        w, w_t, _ = st.weights(None, normalize=True)
        nset = 8
        xs = [torch.randn(N, H, device=dev) for _ in range(nset)]
        ys = [torch.empty(N, H, device=dev) for _ in range(nset)]
        L = lib()

        def k2(i):
            L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xs[i % nset]), H, _p(ys[i % nset]), H, None, N, H, 1,
                   _stream())
        row("K2 SpMM h = 300 + ReLU epilogue (a2)", graph_time(k2), 4 * H * 2 * N + 8 * d.num_items + 4 * (N + 1),
            "32-row CTAs (20 warps/SM, no wave cliff); latency-bound chain: rowptr -> col -> 4 feature batches per warp")
        xf = ops.cast_i64_f32(x_raw)

        def k3(i):
            L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xf), F, _p(ys[i % nset]), F, None, N, F, 0, _stream())
        row("K3 SpMM at input width 9 (GraphConv, a3)", graph_time(k3), 4 * F * 2 * N + 8 * d.num_items + 4 * (N + 1))

        def k4(i):
            torch.ops.ghscn.segment_reduce(xs[i % nset], seg.ptr, None, True)
        row("K4 segment mean [N,300] -> [B,300] (a10)", graph_time(k4), 4 * H * (N + B))
        clusters = (torch.arange(N, device=dev) % K).int()
        hb = hetero.build_hetero_batch(x_raw, ei, batch, clusters, K, padded=True, num_graphs=B)
        lv = hb["local", "to", "virtual"].edge_index
        V = hb["virtual"].x.size(0)
        gat = pyg.GATConv((-1, -1), H, add_self_loops=False).to(dev)
        xv = torch.randn(V, H, device=dev)
        with torch.no_grad():
            gat((xs[0], xv), lv)

            def k5(i):
                gat((xs[i % nset], xv), lv)
            row("K5 GAT cluster pool l->v at width 300, forward (a9)", graph_time(k5), 4 * H * N + 12 * N + 4 * H * V * 2,
                "round-1 chain (row_dot x2, scores, long-row SpMM) + [V,300]x[300,300] projection")
            u = torch.randn(2, H, device=dev)
            pooled = torch.empty(V, H, device=dev)
            lvd = structure_cache().graph(lv, N, V, False).by_dst

            def k5f(i):
                L.call("ghscn_gat_pool_fused_fwd", _p(lvd.rowptr), _p(lvd.col), _p(xs[i % nset]), H, _p(xv), H, _p(u[0]),
                       _p(u[1]), 0.2, V, H, 1, _p(pooled), H, ops.pool_warps_per_row(lvd.col.numel(), V), _stream())
            row("K5 fused attention pool at input width (one pass, online softmax)", graph_time(k5f),
                4 * H * N + 4 * N + 4 * H * V * 2, "what HeteroConv's fused virtual destination launches")
        s_log = torch.randn(N, K, device=dev)
        ei1, _ = pyg.gcn_norm(ei, None, N, add_self_loops=True)

        def k6f(i):
            pyg.mincut_pool_ragged(xs[i % nset], ei1, s_log, batch, want_out=False, want_adj=False)
        E1 = ei1.size(1)
        row("K6 MinCUT losses forward, K = 10 (a4/a5; what hscn.py:63 keeps)", graph_time(k6f),
            4 * N * K * 2 + 4 * (N + 1) + 4 * E1 + B * (8 * K * K + 32), "latency-bound: one CTA (512 threads) per graph, ~10 barrier phases")

        def k6o(i):
            pyg.mincut_pool_ragged(xs[i % nset], ei1, s_log, batch)
        row("K6 MinCUT forward with pooled features and adjacency, K = 10, H = 300", graph_time(k6o),
            4 * N * K * 2 + 4 * N * H + 4 * (N + 1) + 4 * E1 + B * (4 * K * H + 8 * K * K + 32),
            "fused kernel + S^T X streamed by a cluster of CTAs per graph through a TMA bulk-copy ring")
        s_soft = torch.softmax(s_log, -1)

        def k7(i):
            cl = hetero.assign_clusters(s_soft)
            hetero.build_hetero_batch(x_raw, ei, batch, cl, K, padded=True, num_graphs=B, x_float=xf)
        row("K7 argmax + virtual nodes + l->v / v->v edges + their 4 CSRs (a6/a7)", graph_time(k7),
            4 * N * K + 4 * N + 8 * F * N + 4 * F * B * K + 16 * N + 16 * N, "5 launches + torch glue")
        scn = models.SCN([16], "elu", 9, K).to(dev)
        ei_s, ew_s = pyg.gcn_norm(ei, None, N, add_self_loops=True)
        with torch.no_grad():
            def kscn(i):
                scn.logits(xf, ei_s, ew_s)
            row("SCN node pipeline (GraphConv 9->16 + ELU + Linear->10), logits only", graph_time(kscn),
                4 * N * (F + K) + 8 * d.num_items + 4 * (N + 1))
        wgt = torch.randn(H, H, device=dev) / H ** 0.5
        img = gemm.gemm3x_prep(wgt)

        def kg(i):
            gemm.gemm3x(xs[i % nset], img, H, None, False, out=ys[i % nset])
        us = graph_time(kg)
        row("projection x W^T, h x h, tcgen05 3xTF32 (`gemm3x`)", us, 4 * N * 2 * H,
            f"{3 * 2 * N * H * H / us / 1e6:.0f} TFLOP/s of tf32 MMA work; smem-bandwidth-bound main loop (DESIGN 4.1)")

        def ktn(i):
            gemm.gemm3x_tn(ys[i % nset], xs[i % nset])
        us = graph_time(ktn)
        row("weight gradient dY^T x, h x h, tcgen05 3xTF32 (`gemm3x_tn` + slab reduce)", us, 4 * N * 2 * H,
            f"{3 * 2 * N * H * H / us / 1e6:.0f} TFLOP/s of tf32 MMA work")
        w9 = torch.randn(H, F, device=dev)

        def ksk(i):
            gemm.linear(xf, w9, None)
        row("projection 9 -> 300 (`skinny_fwd_reg`)", graph_time(ksk), 4 * N * (F + H))
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
    emit(f"| #1 MPNN(GCN) L=5 h=300 fwd+bwd+AdamW, dropout 0 | 128 | {us / 1e3:.3f} | {128 / us * 1e6:.0f} |")
    structure_cache().clear()
    # config #1, throughput variant: dropout 0.2 (model/mpnn.py:57-58) through the fused ReLU+dropout kernel
    torch.manual_seed(0)
    m = models.MPNN("gcn", torch.relu, 9, 300, 10, 5, dropout=0.2).to(dev)
    m.train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=5e-4, fused=True, capturable=True)
    with structure_hints(**hints):
        for i in range(3):
            step(i)
        us = graph_time(step, reps=5)
    emit(f"| #1 MPNN(GCN) L=5 h=300 fwd+bwd+AdamW, dropout 0.2 (fused ReLU+dropout) | 128 | {us / 1e3:.3f} | {128 / us * 1e6:.0f} |")
    structure_cache().clear()
    # config #3 shape on one GPU
    from graph_hscn_b200.train import GraphHSCNStep, StepConfig
    hb = synthetic.peptides_batch(1024, seed=1237, task="struct")
    from graph_hscn_b200.train import BucketPolicy
    st = GraphHSCNStep(StepConfig(num_classes=11, loss_fn="l1"), hb, dev, padded=True,
                       policy=BucketPolicy(444, 1024, 1024, 1024))
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
    emit(f"# Round-2 sweeps on one B200 (measured HBM copy peak {PEAK:.0f} GB/s)\n")
    which = os.environ.get("SWEEPS", "kernels,spmm,mincut,vocsp,models").split(",")
    t0 = time.time()
    if "kernels" in which:
        kernel_table()
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
