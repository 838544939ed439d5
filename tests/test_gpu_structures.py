"""GPU tests of the sort-free structure shortcuts: they must equal the general stable radix sort bit for bit."""
import pytest
import torch

from tests.util import assert_close, random_edge_index

pytestmark = pytest.mark.gpu


def _same(a, b, nnz=None):
    assert torch.equal(a.rowptr, b.rowptr)
    n = int(a.rowptr[-1]) if nnz is None else nnz
    assert torch.equal(a.col[:n], b.col[:n]) and torch.equal(a.perm[:n], b.perm[:n])


def test_csr_add_loops_equals_sort_with_loops(cuda):
    from graph_hscn_b200.structure import GraphStructure, build_csr
    g = torch.Generator().manual_seed(3)
    n, e = 3000, 9000
    ei = random_edge_index(n, n, e, g, self_loops=False).to(cuda)
    for key, other in ((ei[1], ei[0]), (ei[0], ei[1])):
        full = build_csr(key, other, n, True)
        derived = GraphStructure._with_loops(build_csr(key, other, n, False))
        _same(derived, full)
        assert derived.num_items == ei.size(1) + n


def test_cache_derives_loops_only_when_loop_free(cuda):
    from graph_hscn_b200.structure import StructureCache
    g = torch.Generator().manual_seed(4)
    cache = StructureCache()
    clean = random_edge_index(50, 50, 200, g, self_loops=False).to(cuda)
    dirty = torch.cat([clean, torch.tensor([[3, 7], [3, 7]], device=cuda)], 1)
    assert cache.graph(clean, 50, 50, True)._plain is not None
    assert cache.graph(dirty, 50, 50, True)._plain is None          # existing loops: general path
    d = cache.graph(dirty, 50, 50, True).by_dst
    assert int(d.rowptr[-1]) == clean.size(1) + 50                   # 2 dropped, 50 appended


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("K", [3, 10])
def test_virtual_csr_equals_sort(cuda, padded, K):
    from graph_hscn_b200 import hetero, synthetic
    from graph_hscn_b200.structure import build_csr, structure_cache
    b = synthetic.peptides_batch(9, seed=K)
    g = torch.Generator().manual_seed(K)
    clusters = torch.randint(0, K, (b.x.size(0),), generator=g).int()
    clusters[: int(b.ptr[1])] = 1                      # a graph with a single cluster
    hb = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters.to(cuda), K,
                                   padded=padded)
    N, V = b.x.size(0), hb["virtual"].x.size(0)
    lv = hb["local", "to", "virtual"].edge_index
    vv = hb["virtual", "to", "virtual"].edge_index
    st_lv = structure_cache().graph(lv, N, V, False)
    st_vv = structure_cache().graph(vv, V, V, False)
    assert st_lv._by_dst is not None and st_vv._by_src is not None      # pre-registered, not sorted
    _same(st_lv.by_dst, build_csr(lv[1], lv[0], V, False))
    _same(st_lv.by_src, build_csr(lv[0], lv[1], N, False))
    _same(st_vv.by_dst, build_csr(vv[1], vv[0], V, False))
    _same(st_vv.by_src, build_csr(vv[0], vv[1], V, False))


def test_linear_3xtf32_is_fp32_accurate(cuda):
    """The 3xTF32 tensor-core projection must be as accurate as a plain fp32 GEMM (parity bar 1e-5)."""
    from graph_hscn_b200 import gemm
    from tests.util import rel_err
    g = torch.Generator().manual_seed(6)
    n, k, m = 4096, 300, 300
    x = (torch.randn(n, k, generator=g) * 3).to(cuda).requires_grad_()
    w = (torch.randn(m, k, generator=g) / 17).to(cuda).requires_grad_()
    b = torch.randn(m, generator=g).to(cuda).requires_grad_()
    gy = torch.randn(n, m, generator=g).to(cuda)
    hi, lo = gemm.split_tf32(x.detach())
    assert torch.equal(hi + lo, x.detach()) and bool(((hi.view(torch.int32) & 0x1FFF) == 0).all())
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), gy.double())
    gemm.set_gemm_mode("3xtf32")
    y = gemm.linear(x, w, b)
    dx, dw, db = torch.autograd.grad(y, (x, w, b), gy)
    gemm.set_gemm_mode("fp32")
    y32 = gemm.linear(x, w, b)
    dx32, dw32, db32 = torch.autograd.grad(y32, (x, w, b), gy)
    gemm.set_gemm_mode("3xtf32")
    for name, a, a32, r in [("y", y, y32, ref), ("dx", dx, dx32, gx), ("dw", dw, dw32, gw), ("db", db, db32, gb)]:
        e3, e32 = rel_err(a, r.float()), rel_err(a32, r.float())
        bar = 5e-6                             # single-pass tcgen05 accumulation (~3e-6); 2x inside the 1e-5 bar
        assert e3 < bar, f"{name}: 3xTF32 error {e3:.2e} (plain fp32 {e32:.2e})"
    torch.backends.cuda.matmul.allow_tf32 = True          # plain TF32 would NOT meet the bar
    ytf = torch.nn.functional.linear(x, w, b)
    torch.backends.cuda.matmul.allow_tf32 = False
    assert rel_err(ytf, ref.float()) > 1e-5


def test_gcn_stack_at_bench_width_meets_parity_bar(cuda):
    """3 GCN layers at h=300 on >4096 nodes (the 3xTF32 projection path is active) vs the CPU oracle."""
    from graph_hscn_b200 import gemm, pyg, synthetic
    from oracle.namespace import namespace
    from tests.util import RTOL, rel_err
    o, p = namespace(), pyg.namespace()
    b = synthetic.peptides_batch(40, seed=90)
    n = b.x.size(0)
    assert n >= gemm.MIN_ROWS and gemm.gemm_mode() == "3xtf32"
    torch.manual_seed(3)
    ref = [o.GCNConv(9, 300, add_self_loops=False), o.GCNConv(300, 300, add_self_loops=False),
           o.GCNConv(300, 300, add_self_loops=False)]
    tst = [p.GCNConv(9, 300, add_self_loops=False).to(cuda) for _ in range(1)] + \
          [p.GCNConv(300, 300, add_self_loops=False).to(cuda) for _ in range(2)]
    for r, t in zip(ref, tst):
        t.load_state_dict(r.state_dict())
    xr = b.x.float()
    xt = xr.to(cuda)
    ei_t = b.edge_index.to(cuda)
    for r, t in zip(ref, tst):
        xr, xt = torch.relu(r(xr, b.edge_index)), torch.relu(t(xt, ei_t))
    assert_close(xt, xr, RTOL, "3-layer GCN features at h=300")
    xr.sum().backward()
    xt.sum().backward()
    for i, (r, t) in enumerate(zip(ref, tst)):
        assert rel_err(t.lin.weight.grad, r.lin.weight.grad) < 5 * RTOL, f"dW layer {i}"
        assert rel_err(t.bias.grad, r.bias.grad) < 5 * RTOL, f"db layer {i}"


def test_colsum_matches_torch(cuda):
    g = torch.Generator().manual_seed(5)
    for n, f in [(1, 4), (18269, 300), (1000, 10), (65, 33)]:
        x = torch.randn(n, f, generator=g).to(cuda)
        assert_close(torch.ops.ghscn.colsum(x), x.double().sum(0).float(), 1e-5, f"colsum {n}x{f}")


@pytest.mark.parametrize("k,m", [(9, 300), (9, 16), (16, 10), (32, 64)])
def test_skinny_linear_fwd_bwd(cuda, k, m):
    """Hand-written tall-skinny projection kernels vs an fp64 reference (and they must be the active path)."""
    from graph_hscn_b200 import gemm
    from tests.util import rel_err
    g = torch.Generator().manual_seed(k * m)
    n = 5000
    x = torch.randint(0, 12, (n, k), generator=g).float().to(cuda).requires_grad_()
    w = (torch.randn(m, k, generator=g) / 3).to(cuda).requires_grad_()
    b = torch.randn(m, generator=g).to(cuda).requires_grad_()
    gy = torch.randn(n, m, generator=g).to(cuda)
    y = gemm.linear(x, w, b)
    assert type(y.grad_fn).__name__.startswith("_LinearSkinny")
    dx, dw, db = torch.autograd.grad(y, (x, w, b), gy)
    ref = torch.nn.functional.linear(x.double(), w.double(), b.double())
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), gy.double())
    for name, a, r in [("y", y, ref), ("dx", dx, gx), ("dw", dw, gw), ("db", db, gb)]:
        assert rel_err(a, r.float()) < 2e-6, name


def _collated(sizes, degree, gen, shuffle_within=True):
    """Collated edge list (graph-major, block diagonal) of random graphs with the given node counts."""
    parts, batch, off = [], [], 0
    for g, n in enumerate(sizes):
        e = int(n * degree)
        if n > 0 and e > 0:
            ei = torch.randint(0, n, (2, e), generator=gen)
            if not shuffle_within:
                ei = ei[:, torch.argsort(ei[1], stable=True)]
            parts.append(ei + off)
        batch += [g] * n
        off += n
    return torch.cat(parts, 1), torch.tensor(batch, dtype=torch.long)


@pytest.mark.parametrize("case", ["peptides", "ragged", "dense_rows", "voc"])
def test_blocked_csr_equals_radix_sort(cuda, case):
    """K1 fast path (one CTA per graph, both orientations) == two stable radix sorts, bit for bit."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.structure import StructureCache, build_csr, edge_blocks_from_batch
    g = torch.Generator().manual_seed(11)
    if case == "peptides":
        b = synthetic.peptides_batch(128, seed=5)
        ei, batch = b.edge_index, b.batch
    elif case == "ragged":          # empty graphs, single nodes, edge-free graphs, duplicates and self loops
        ei, batch = _collated([5, 0, 1, 300, 0, 2, 64, 1, 444, 3], 2.5, g)
        keep = (batch[ei[0]] != 6)                       # graph 6 keeps its nodes but loses every edge
        ei = ei[:, keep]
    elif case == "dense_rows":      # stars: one destination collects hundreds of edges (rank loop, long rows)
        ei, batch = _collated([200, 333, 50], 6.0, g)
        ei[1, ei[1] < 200] = 7
    else:
        ei, batch = _collated([int(v) for v in torch.randint(395, 501, (32,), generator=g)], 5.66, g)
    B = int(batch.max()) + 1
    N = batch.numel()
    blocks = edge_blocks_from_batch(ei, batch, B)
    assert blocks is not None
    ei_d, batch_d = ei.to(cuda), batch.to(cuda)
    cache = StructureCache()
    seg = cache.segments(batch_d, B)
    cache.register_blocks(ei_d, seg.ptr, B, blocks[0], blocks[1])
    st = cache.graph(ei_d, N, N, False)
    assert st._blocks is not None
    _same(st.by_dst, build_csr(ei_d[1], ei_d[0], N, False))
    _same(st.by_src, build_csr(ei_d[0], ei_d[1], N, False))
    assert int(cache.blocked_status(cuda)) == 0
    assert int(st.by_dst.rowptr[-1]) == ei.size(1)


def test_blocked_csr_reports_broken_promises(cuda):
    from graph_hscn_b200.structure import StructureCache, edge_blocks_from_batch
    g = torch.Generator().manual_seed(12)
    ei, batch = _collated([40, 50, 60], 3.0, g)
    bad = ei.clone()
    bad[1, 5] = 100                                       # an edge of graph 0 pointing into graph 2
    assert edge_blocks_from_batch(bad, batch, 3) is None  # the host-side check refuses it
    assert edge_blocks_from_batch(ei.flip(1), batch, 3) is None          # not graph-major
    cache = StructureCache()
    seg = cache.segments(batch.to(cuda), 3)
    cache.blocked_status(cuda).zero_()
    bad_d = bad.to(cuda)
    cache.register_blocks(bad_d, seg.ptr, 3, 60, 180)                    # a false promise: the kernel flags it
    cache.graph(bad_d, 150, 150, False).by_dst
    assert int(cache.blocked_status(cuda)) & 2
    cache.blocked_status(cuda).zero_()
    ok = ei.to(cuda)
    cache.register_blocks(ok, seg.ptr, 3, 60, 10)                        # bound too small
    cache.graph(ok, 150, 150, False).by_dst
    assert int(cache.blocked_status(cuda)) & 1
    cache.blocked_status(cuda).zero_()


def test_batch_to_device_registers_the_fast_csr_path(cuda):
    """A collated Batch moved to the GPU builds its CSRs with the per-graph kernel; the convs see the same numbers."""
    from graph_hscn_b200 import pyg, synthetic
    from graph_hscn_b200.structure import build_csr, structure_cache
    structure_cache().clear()
    structure_cache().blocked_status(cuda).zero_()
    b = synthetic.peptides_batch(20, seed=41).to(cuda)
    N = b.x.size(0)
    st = structure_cache().graph(b.edge_index, N, N, False)
    assert st._blocks is not None and st._blocks.max_edges == b.max_edges_per_graph
    _same(st.by_dst, build_csr(b.edge_index[1], b.edge_index[0], N, False))
    _same(st.by_src, build_csr(b.edge_index[0], b.edge_index[1], N, False))
    assert int(structure_cache().blocked_status(cuda)) == 0
    torch.manual_seed(0)
    conv = pyg.GCNConv(9, 32).to(cuda)
    y_fast = conv(b.x.float(), b.edge_index)
    ei_plain = b.edge_index.clone()                      # an unregistered copy of the same edges: radix path
    assert structure_cache().graph(ei_plain, N, N, True)._plain._blocks is None
    assert torch.equal(conv(b.x.float(), ei_plain), y_fast)


@pytest.mark.parametrize("rows,k,m,act", [(129, 300, 300, "relu"), (129, 300, 10, None), (1290, 600, 300, "relu"),
                                          (1027, 300, 300, "elu"), (7, 5, 3, "tanh"), (300, 18, 300, None)])
def test_small_linear_fwd_bwd(cuda, rows, k, m, act):
    """csrc/dense_small.cu (graph-level head, virtual-node projections) vs an fp64 torch reference: output, dx, dW, db
    with the activation folded into the kernels."""
    import torch.nn.functional as F
    from graph_hscn_b200 import gemm
    g = torch.Generator().manual_seed(rows + k)
    x = torch.randn(rows, k, generator=g)
    w = torch.randn(m, k, generator=g) / k ** 0.5
    b = torch.randn(m, generator=g)
    gy = torch.randn(rows, m, generator=g)
    fn = {"relu": F.relu, "elu": F.elu, "tanh": torch.tanh, None: lambda t: t}[act]
    xr, wr, br = (t.double().requires_grad_() for t in (x, w, b))
    yr = fn(F.linear(xr, wr, br))
    yr.backward(gy.double())
    xt, wt, bt = (t.to(cuda).requires_grad_() for t in (x, w, b))
    yt = gemm.linear_act(xt, wt, bt, act)
    assert yt is not None
    yt.backward(gy.to(cuda))
    assert_close(yt, yr.float(), 2e-6, "small linear out")
    assert_close(xt.grad, xr.grad.float(), 2e-6, "small linear dx")
    assert_close(wt.grad, wr.grad.float(), 2e-6, "small linear dW")
    assert_close(bt.grad, br.grad.float(), 2e-6, "small linear db")
    # the plain dispatcher takes the same kernels for < 4 k rows
    y2 = gemm.linear(x.to(cuda), w.to(cuda), b.to(cuda))
    assert_close(y2, F.linear(x.double(), w.double(), b.double()).float(), 2e-6, "gemm.linear small path")


@pytest.mark.parametrize("loss_fn,rows,total,c", [("cross_entropy", 128, 129, 10), ("l1", 1024, 1027, 11),
                                                  ("cross_entropy", 5, 5, 3)])
def test_graph_loss_matches_criterion(cuda, loss_fn, rows, total, c):
    """Fused task loss (loss.py:6-19): value, sigmoid score and gradient vs models.criterion on the first `rows` rows;
    padding rows get zero gradient."""
    from graph_hscn_b200 import models, ops
    g = torch.Generator().manual_seed(total)
    pred = torch.randn(total, c, generator=g) * 3
    y = (torch.rand(total, c, generator=g) < 0.3).float() if loss_fn == "cross_entropy" else torch.randn(total, c, generator=g)
    pr = pred.clone().requires_grad_()
    lr, sr = models.criterion(loss_fn, pr[:rows], y[:rows])
    (lr * 1.7).backward()
    pt = pred.to(cuda).requires_grad_()
    lt, st_ = ops.graph_loss(loss_fn, pt, y.to(cuda), rows=rows)
    (lt * 1.7).backward()
    assert_close(lt, lr, 1e-6, "loss")
    assert_close(st_, sr, 1e-6, "score")
    assert_close(pt.grad, pr.grad, 1e-6, "d pred")
    assert not pt.grad[rows:].any()


@pytest.mark.parametrize("loss_fn,rows,total,h,c,bias,unit", [("cross_entropy", 128, 130, 300, 10, True, False),
                                                             ("l1", 200, 256, 128, 11, True, True),
                                                             ("cross_entropy", 5, 5, 512, 3, False, False)])
def test_head_out_loss_matches_unfused(cuda, loss_fn, rows, total, h, c, bias, unit):
    """Output layer + criterion + their backward in one launch (model/hscn.py:112, loss.py:6-19) vs the float64
    composition F.linear -> criterion -> autograd; padding rows get no gradient."""
    import torch.nn.functional as F
    from graph_hscn_b200 import models, ops
    g = torch.Generator().manual_seed(total + c)
    hid = torch.randn(total, h, generator=g)
    w = torch.randn(c, h, generator=g) / h ** 0.5
    b = torch.randn(c, generator=g) if bias else None
    y = (torch.rand(total, c, generator=g) < 0.3).float() if loss_fn == "cross_entropy" else torch.randn(total, c, generator=g)
    scale = 1.0 if unit else 1.7
    ref = [t.double().requires_grad_() if t is not None else None for t in (hid, w, b)]
    pred_r = F.linear(ref[0], ref[1], ref[2])
    lr, sr = models.criterion(loss_fn, pred_r[:rows], y[:rows].double())
    (lr * scale).backward()
    dev = [t.to(cuda).requires_grad_() if t is not None else None for t in (hid, w, b)]
    assert ops.head_out_loss_ok(total, h, c)
    lt, pred, score = ops.head_out_loss(loss_fn, dev[0], dev[1], dev[2], y.to(cuda), rows=rows, unit_grad=unit)
    (lt * scale).backward()
    assert_close(lt, lr.float(), 2e-6, "loss")
    assert_close(pred, pred_r.float(), 2e-6, "pred")
    assert_close(score, sr.float(), 2e-6, "score")
    for name, a, r in zip(("d hidden", "d W2", "d b2"), dev, ref):
        if a is not None:
            assert_close(a.grad, r.grad.float(), 2e-6, name)
    assert not dev[0].grad[rows:].any()
    assert not ops.head_out_loss_ok(1027, 300, 11)                  # larger batches keep the separate kernels


@pytest.mark.parametrize("n,f", [(5000, 300), (777, 37), (1, 4)])
def test_relu_grad_colsum_matches_threshold_and_sum(cuda, n, f):
    """ReLU backward + first stage of the bias gradient in one pass (model/hscn.py:110 `.relu()` behind GCNConv): the
    masked gradient is bit-identical to threshold_backward, the column sums match float64."""
    from graph_hscn_b200 import ops
    g = torch.Generator().manual_seed(n)
    dy = torch.randn(n, f, generator=g).to(cuda)
    y = torch.relu(torch.randn(n, f, generator=g)).to(cuda)
    for fork in (True, False):
        out, db, join = ops.relu_grad_colsum(dy, y, fork=fork)
        join()
        want = torch.ops.aten.threshold_backward(dy, y, 0.0)
        assert torch.equal(out, want)
        assert_close(db, want.double().sum(0).float(), 2e-6, "db")


@pytest.mark.parametrize("feat", [9, 300])
def test_attention_pool_warps_per_row_agree(cuda, feat):
    """The fused attention pool (model/hscn.py:88 GATConv local -> virtual) gives the same softmax-weighted sums whether
    a destination row is pooled by 1, 2, 4 or 8 warps (partials merged by log-sum-exp), for empty, short and long rows,
    and matches a float64 evaluation of the same formula."""
    from graph_hscn_b200._lib import lib
    from graph_hscn_b200.structure import _p, _stream
    g = torch.Generator().manual_seed(feat)
    lens = torch.tensor([0, 1, 2, 5, 15, 16, 31, 32, 33, 70, 200] * 3)
    V, N = lens.numel(), int(lens.sum())
    rowptr = torch.zeros(V + 1, dtype=torch.int32)
    rowptr[1:] = lens.cumsum(0)
    col = torch.randperm(N, generator=g).int()
    xs, xd = torch.randn(N, feat, generator=g), torch.randn(V, feat, generator=g)
    u = torch.randn(2, feat, generator=g) / feat ** 0.5
    want = torch.zeros(V, feat, dtype=torch.float64)
    for r in range(V):
        m = col[rowptr[r]:rowptr[r + 1]].long()
        if m.numel():
            z = torch.nn.functional.leaky_relu(xs[m].double() @ u[0].double() + xd[r].double() @ u[1].double(), 0.2)
            a = torch.softmax(z, 0)
            want[r] = (a[:, None] * xs[m].double()).sum(0)
    dev = [t.to(cuda) for t in (rowptr, col, xs, xd, u)]
    outs = []
    for wpr in (1, 2, 4, 8):
        pooled = torch.full((V, feat), float("nan"), device=cuda)
        lib().call("ghscn_gat_pool_fused_fwd", _p(dev[0]), _p(dev[1]), _p(dev[2]), feat, _p(dev[3]), feat, _p(dev[4][0]),
                   _p(dev[4][1]), 0.2, V, feat, 1, _p(pooled), feat, wpr, _stream())
        assert_close(pooled, want.float(), 2e-6, f"pool, {wpr} warps per row")
        outs.append(pooled)
    assert not outs[0][lens == 0].any()                             # empty rows pool to zero


def test_relu_dropout_fused(cuda):
    """model/mpnn.py:57-58 `F.dropout(self.activation(x), p)` as one kernel: keep rate, scaling, zeros where relu is
    zero, a fresh mask per call (also across CUDA-graph replays), backward from the output alone."""
    from graph_hscn_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4000, 300, generator=g).to(cuda).requires_grad_()
    p = 0.2
    y = ops.relu_dropout(x, p, True)
    pos = x.detach() > 0
    kept = y.detach() > 0
    assert not kept[~pos].any()                                              # relu zeros stay zero
    rate = float(kept[pos].float().mean())
    assert abs(rate - (1 - p)) < 5e-3, rate                                  # ~480k positive elements
    assert torch.allclose(y.detach()[kept], x.detach()[kept] / (1 - p), rtol=1e-6)
    gy = torch.randn(y.shape, generator=g).to(cuda)
    y.backward(gy)
    want = torch.where(kept, gy / (1 - p), torch.zeros_like(gy))
    assert torch.allclose(x.grad, want, rtol=1e-6)
    y2 = ops.relu_dropout(x.detach(), p, True)
    assert float(((y2 > 0) != kept).float().mean()) > 0.05                   # a different mask on the next call
    assert torch.equal(ops.relu_dropout(x.detach(), p, False), torch.relu(x.detach()))
    # replays of a captured graph draw new masks (the call counter lives on the device)
    xs = x.detach()
    out = torch.empty_like(xs)
    ops.relu_dropout(xs, p, True)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out.copy_(ops.relu_dropout(xs, p, True))
    graph.replay()
    a = out.clone()
    graph.replay()
    torch.cuda.synchronize()
    assert float(((a > 0) != (out > 0)).float().mean()) > 0.05


def test_mpnn_with_dropout_trains_on_the_fused_path(cuda):
    """Config #1's throughput variant (dropout 0.2): forward + backward run, outputs are finite, eval mode is exact."""
    import torch.nn.functional as F
    from graph_hscn_b200 import models, pyg, synthetic
    b = synthetic.peptides_batch(8, seed=9).to(cuda)
    b.x = b.x.float()
    torch.manual_seed(0)
    m = models.MPNN("gcn", F.relu, 9, 64, 10, 4, dropout=0.2, ops=pyg.namespace()).to(cuda)
    m.train()
    out = m(b)
    loss, _ = models.criterion("cross_entropy", out, b.y)
    loss.backward()
    assert torch.isfinite(out).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    o1, o2 = m(b), m(b)
    assert torch.equal(o1, o2)
