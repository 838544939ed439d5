"""GPU parity tests: CUDA path (through the C ABI / torch custom ops) vs the CPU oracle.

Bars (BASELINE.json north_star): bit-exact for CSR construction, cluster argmax and virtual-node
indexing; <= 1e-5 relative (fp32) for features, pooled adjacency, losses and gradients.
"""
import numpy as np
import pytest
import torch

from tests.util import RTOL, assert_close, random_edge_index, rel_err

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle.namespace import namespace
    return namespace()


def _product():
    from graph_hscn_b200 import pyg
    return pyg.namespace()


def _to_dev(model_cpu_ctor, model_gpu_ctor, dev, prime):
    """Build the oracle model on CPU and the product model on GPU with identical parameters."""
    torch.manual_seed(0)
    ref = model_cpu_ctor()
    prime(ref, "cpu")               # materialise lazy parameters
    tst = model_gpu_ctor().to(dev)
    from torch.nn.parameter import UninitializedParameter
    sd = ref.state_dict()
    for name, p in tst.named_parameters():
        if isinstance(p, UninitializedParameter):
            p.materialize(sd[name].shape, device=dev)
    tst.load_state_dict(sd)
    return ref, tst


# ------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (7, 5, 1), (300, 2000, 2), (19200, 39296, 3), (70000, 300000, 4),
                                       (5, 4097, 5)])
@pytest.mark.parametrize("loops", [False, True])
def test_csr_build_bit_exact(cuda, n, e, seed, loops):
    from graph_hscn_b200.structure import build_csr
    from oracle import ops as oops
    g = torch.Generator().manual_seed(seed)
    ei = random_edge_index(n, n, e, g)
    if loops:
        ei2, _ = oops.add_remaining_self_loops(ei, None, 1.0, n)
        keep_ids = torch.cat([torch.nonzero(ei[0] != ei[1]).flatten(), e + torch.arange(n)])
    else:
        ei2, keep_ids = ei, torch.arange(e)
    order = torch.sort(ei2[1], stable=True)[1]
    want_perm = keep_ids[order].int()
    want_col = ei2[0][order].int()
    want_rowptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(ei2[1], minlength=n).cumsum(0)]).int()

    eid = ei.to(cuda)
    csr = build_csr(eid[1], eid[0], n, loops)
    nnz = ei2.size(1)
    assert torch.equal(csr.rowptr.cpu(), want_rowptr)
    assert torch.equal(csr.perm[:nnz].cpu(), want_perm)
    assert torch.equal(csr.col[:nnz].cpu(), want_col)
    if nnz < csr.num_items:      # dropped loops: tail holds them in edge order, col = -1
        assert bool((csr.col[nnz:] == -1).all())
        dropped = torch.nonzero(ei[0] == ei[1]).flatten().int()
        assert torch.equal(csr.perm[nnz:].cpu(), dropped)


def test_csr_ignores_negative_padding(cuda):
    from graph_hscn_b200.structure import build_csr
    ei = torch.tensor([[0, -1, 2, 1, -1], [1, -1, 0, 1, -1]])
    csr = build_csr(ei[1].to(cuda), ei[0].to(cuda), 3, False)
    assert csr.rowptr.tolist() == [0, 1, 3, 3]
    assert csr.perm[:3].tolist() == [2, 0, 3]


def test_batch_to_ptr(cuda):
    from graph_hscn_b200.structure import structure_cache
    batch = torch.tensor([0, 0, 0, 2, 2, 5])
    seg = structure_cache().segments(batch.to(cuda), 7)
    assert seg.ptr.tolist() == [0, 3, 3, 5, 5, 5, 6, 6]


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("loops", [False, True])
def test_gcn_norm_bit_exact(cuda, weighted, loops):
    from graph_hscn_b200.pyg import gcn_norm
    from oracle import ops as oops
    g = torch.Generator().manual_seed(11)
    n, e = 500, 3000
    ei = random_edge_index(n, n, e, g)
    w = torch.rand(e, generator=g) + 0.1 if weighted else None
    ref_ei, ref_w = oops.gcn_norm(ei, w, n, False, loops)
    got_ei, got_w = gcn_norm(ei.to(cuda), None if w is None else w.to(cuda), n, False, loops)
    assert torch.equal(got_ei.cpu(), ref_ei)
    assert torch.equal(got_w.cpu(), ref_w), f"max diff {(got_w.cpu() - ref_w).abs().max()}"


# ------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("feat", [1, 9, 10, 16, 64, 300, 512, 700])
def test_spmm_matches_scatter_add(cuda, feat):
    from graph_hscn_b200.structure import structure_cache
    from oracle import ops as oops
    g = torch.Generator().manual_seed(feat)
    n, e = 1000, 4000
    ei = random_edge_index(n, n, e, g)
    w = torch.randn(e, generator=g)
    x = torch.randn(n, feat, generator=g)
    ref = oops.propagate_add(x, ei, w, n)
    st = structure_cache().graph(ei.to(cuda), n, n, False)
    wd, wt, _ = st.weights(w.to(cuda), normalize=False)
    d, s = st.by_dst, st.by_src
    got = torch.ops.ghscn.spmm(d.rowptr, d.col, wd, s.rowptr, s.col, wt, x.to(cuda), None)
    # same summation order and roundings as CPU scatter_add_: expect (near) bit equality
    assert_close(got, ref, 1e-6, f"spmm F={feat}")
    exact = float((got.cpu() == ref).float().mean())
    assert exact > 0.99, f"only {exact:.3f} of outputs bit-equal"


@pytest.mark.parametrize("add_loops", [True, False])
def test_gcnconv_fwd_bwd(cuda, add_loops):
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(5)
    n, e, fi, fo = 800, 2400, 9, 300
    ei = random_edge_index(n, n, e, g)
    x = torch.randn(n, fi, generator=g)
    ref, tst = _to_dev(lambda: o.GCNConv(fi, fo, add_self_loops=add_loops),
                       lambda: p.GCNConv(fi, fo, add_self_loops=add_loops), cuda, lambda m, d: None)
    with torch.no_grad():
        ref.bias.uniform_(-1, 1)
        tst.bias.copy_(ref.bias)
    xr = x.clone().requires_grad_()
    xt = x.to(cuda).requires_grad_()
    yr = ref(xr, ei)
    yt = tst(xt, ei.to(cuda))
    assert_close(yt, yr, RTOL, "GCNConv out")
    gy = torch.randn(n, fo, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xt.grad, xr.grad, RTOL, "GCNConv dx")
    assert_close(tst.lin.weight.grad, ref.lin.weight.grad, RTOL, "GCNConv dW")
    assert_close(tst.bias.grad, ref.bias.grad, RTOL, "GCNConv db")


def test_graphconv_weighted_fwd_bwd(cuda):
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(6)
    n, e = 150, 460
    ei = random_edge_index(n, n, e, g, self_loops=False)
    x = torch.randint(0, 10, (n, 9), generator=g).float()
    ei_n, w_n = o.gcn_norm(ei, None, n, add_self_loops=True)
    ref, tst = _to_dev(lambda: o.GraphConv(9, 16), lambda: p.GraphConv(9, 16), cuda, lambda m, d: None)
    wr = w_n.clone().requires_grad_()
    wt = w_n.to(cuda).requires_grad_()
    yr = ref(x, ei_n, wr)
    yt = tst(x.to(cuda), ei_n.to(cuda), wt)
    assert_close(yt, yr, RTOL, "GraphConv out")
    yr.sum().backward()
    yt.sum().backward()
    assert_close(tst.lin_rel.weight.grad, ref.lin_rel.weight.grad, RTOL, "dW_rel")
    assert_close(tst.lin_root.weight.grad, ref.lin_root.weight.grad, RTOL, "dW_root")
    assert_close(wt.grad, wr.grad, RTOL, "d edge_weight")


# ------------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("feat", [10, 11, 300])
@pytest.mark.parametrize("sorted_index", [True, False])
def test_scatter_mean_fwd_bwd(cuda, feat, sorted_index):
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(7)
    B, n = 37, 3000
    idx = torch.randint(0, B, (n,), generator=g)
    idx[idx == 5] = 6  # an empty segment
    if sorted_index:
        idx = idx.sort()[0]
    x = torch.randn(n, feat, generator=g)
    xr = x.clone().requires_grad_()
    xt = x.to(cuda).requires_grad_()
    yr = o.scatter_mean(xr, idx, dim=0, dim_size=B)
    yt = p.scatter_mean(xt, idx.to(cuda), dim=0, dim_size=B)
    assert_close(yt, yr, 1e-6, "scatter_mean")
    gy = torch.randn(B, feat, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xt.grad, xr.grad, 1e-6, "scatter_mean dx")
    # global_mean_pool with inferred size
    assert_close(p.global_mean_pool(x.to(cuda), idx.to(cuda)), o.global_mean_pool(x, idx), 1e-6, "gmp")


# ------------------------------------------------------------------------------------------- K5
def test_gat_bipartite_fwd_bwd(cuda):
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(8)
    N, V, fi, H = 1500, 100, 9, 300
    lv = torch.stack([torch.arange(N), torch.randint(0, V - 3, (N,), generator=g)])   # last 3 virtual: empty
    xl = torch.randn(N, fi, generator=g)
    xv = torch.randn(V, fi, generator=g)

    def prime(m, d):
        m((xl.to(d), xv.to(d)), lv.to(d))
    ref, tst = _to_dev(lambda: o.GATConv((-1, -1), H, add_self_loops=False),
                       lambda: p.GATConv((-1, -1), H, add_self_loops=False), cuda, prime)
    with torch.no_grad():
        ref.bias.uniform_(-1, 1)
        tst.bias.copy_(ref.bias)
    xlr, xvr = xl.clone().requires_grad_(), xv.clone().requires_grad_()
    xlt, xvt = xl.to(cuda).requires_grad_(), xv.to(cuda).requires_grad_()
    yr = ref((xlr, xvr), lv)
    yt = tst((xlt, xvt), lv.to(cuda))
    assert_close(yt, yr, RTOL, "GAT out")
    gy = torch.randn(V, H, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xlt.grad, xlr.grad, 2 * RTOL, "GAT dx_local")
    assert_close(xvt.grad, xvr.grad, 2 * RTOL, "GAT dx_virtual")
    for name in ["lin_src.weight", "lin_dst.weight", "att_src", "att_dst", "bias"]:
        gr = dict(ref.named_parameters())[name].grad
        gt = dict(tst.named_parameters())[name].grad
        assert_close(gt, gr, 2 * RTOL, f"GAT d{name}")


def test_gat_homogeneous_self_loops(cuda):
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(9)
    n, e = 400, 1500
    ei = random_edge_index(n, n, e, g)
    x = torch.randn(n, 9, generator=g)
    ref, tst = _to_dev(lambda: o.GATConv(9, 32), lambda: p.GATConv(9, 32), cuda, lambda m, d: None)
    assert_close(tst(x.to(cuda), ei.to(cuda)), ref(x, ei), RTOL, "GAT homogeneous")


@pytest.mark.parametrize("concat", [True, False])
def test_gat_multi_head_fwd_bwd(cuda, concat):
    """heads = 3: bipartite and homogeneous forms against the oracle, values and gradients."""
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(10)
    N, V, fi, C, heads = 900, 60, 9, 16, 3
    lv = torch.stack([torch.arange(N), torch.randint(0, V, (N,), generator=g)])
    xl = torch.randn(N, fi, generator=g)
    xv = torch.randn(V, fi, generator=g)
    ref, tst = _to_dev(lambda: o.GATConv((-1, -1), C, heads=heads, concat=concat, add_self_loops=False),
                       lambda: p.GATConv((-1, -1), C, heads=heads, concat=concat, add_self_loops=False), cuda,
                       lambda m, d: m((xl.to(d), xv.to(d)), lv.to(d)))
    xlr, xvr = xl.clone().requires_grad_(), xv.clone().requires_grad_()
    xlt, xvt = xl.to(cuda).requires_grad_(), xv.to(cuda).requires_grad_()
    yr, yt = ref((xlr, xvr), lv), tst((xlt, xvt), lv.to(cuda))
    assert yt.shape == yr.shape == (V, heads * C if concat else C)
    assert_close(yt, yr, RTOL, "multi-head GAT out")
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xlt.grad, xlr.grad, 2 * RTOL, "multi-head GAT dx_local")
    assert_close(xvt.grad, xvr.grad, 2 * RTOL, "multi-head GAT dx_virtual")
    for name in ["lin_src.weight", "lin_dst.weight", "att_src", "att_dst", "bias"]:
        assert_close(dict(tst.named_parameters())[name].grad, dict(ref.named_parameters())[name].grad, 2 * RTOL,
                     f"multi-head GAT d{name}")
    ei = random_edge_index(300, 300, 1200, g)
    x = torch.randn(300, fi, generator=g)
    ref2, tst2 = _to_dev(lambda: o.GATConv(fi, C, heads=heads, concat=concat),
                         lambda: p.GATConv(fi, C, heads=heads, concat=concat), cuda, lambda m, d: None)
    assert_close(tst2(x.to(cuda), ei.to(cuda)), ref2(x, ei), RTOL, "multi-head GAT homogeneous")


# ------------------------------------------------------------------------------------------- K6
def _peptide_batch(num_graphs, seed):
    from graph_hscn_b200 import synthetic
    return synthetic.peptides_batch(num_graphs, seed=seed)


@pytest.mark.parametrize("K,H", [(10, 16), (4, 64), (32, 40), (128, 24), (64, 300), (128, 512), (68, 332),
                                 (10, 300), (5, 132), (6, 256), (16, 128), (31, 512)])   # streamed pooled features, fwd + bwd
def test_mincut_ragged_fwd_bwd(cuda, K, H):
    o, p = _oracle(), _product()
    b = _peptide_batch(6, seed=K)
    g = torch.Generator().manual_seed(K)
    N = b.x.size(0)
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)   # binary A + I, as hscn.py:61 sees it
    x = torch.randn(N, H, generator=g)
    s = torch.randn(N, K, generator=g)
    xr, sr = x.clone().requires_grad_(), s.clone().requires_grad_()
    xt, st_ = x.to(cuda).requires_grad_(), s.to(cuda).requires_grad_()
    out_r, adj_r, mc_r, or_r = o.mincut_pool_ragged(xr, ei, sr, b.batch)
    out_t, adj_t, mc_t, or_t = p.mincut_pool_ragged(xt, ei.to(cuda), st_, b.batch.to(cuda))
    assert_close(out_t, out_r, RTOL, "mincut out")
    assert_close(adj_t, adj_r, RTOL, "mincut out_adj")
    assert_close(mc_t, mc_r, RTOL, "mincut loss")
    assert_close(or_t, or_r, RTOL, "ortho loss")
    # full backward: all four outputs feed the objective
    go = torch.randn(out_r.shape, generator=g)
    ga = torch.randn(adj_r.shape, generator=g)
    (mc_r * 1.3 + or_r * 0.7 + (out_r * go).sum() * 0.01 + (adj_r * ga).sum()).backward()
    (mc_t * 1.3 + or_t * 0.7 + (out_t * go.to(cuda)).sum() * 0.01 + (adj_t * ga.to(cuda)).sum()).backward()
    assert_close(st_.grad, sr.grad, 5 * RTOL, "mincut d logits")
    assert_close(xt.grad, xr.grad, 5 * RTOL, "mincut d x")


def test_mincut_strided_features_fwd_bwd(cuda):
    """x as a column slice of a wider matrix (row stride != H): the pooled-feature forward stays in the per-graph
    kernel, the backward streams the rows with one bulk copy per row."""
    o, p = _oracle(), _product()
    b = _peptide_batch(5, seed=31)
    g = torch.Generator().manual_seed(31)
    N, K, H = b.x.size(0), 10, 128
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    wide = torch.randn(N, H + 8, generator=g)
    s = torch.randn(N, K, generator=g)
    xr, sr = wide[:, :H].clone().requires_grad_(), s.clone().requires_grad_()
    wide_t = wide.to(cuda).requires_grad_()
    xt, st_ = wide_t[:, :H], s.to(cuda).requires_grad_()
    assert xt.stride(0) == H + 8
    out_r, adj_r, mc_r, or_r = o.mincut_pool_ragged(xr, ei, sr, b.batch)
    out_t, adj_t, mc_t, or_t = p.mincut_pool_ragged(xt, ei.to(cuda), st_, b.batch.to(cuda))
    assert_close(out_t, out_r, RTOL, "mincut out (strided x)")
    go = torch.randn(out_r.shape, generator=g)
    (mc_r + or_r + (out_r * go).sum() * 0.01 + adj_r.sum()).backward()
    (mc_t + or_t + (out_t * go.to(cuda)).sum() * 0.01 + adj_t.sum()).backward()
    assert_close(st_.grad, sr.grad, 5 * RTOL, "d logits (strided x)")
    assert_close(wide_t.grad[:, :H], xr.grad, 5 * RTOL, "d x (strided x)")
    assert float(wide_t.grad[:, H:].abs().max()) == 0.0


@pytest.mark.parametrize("K,H", [(10, 16), (64, 16), (128, 300)])
def test_mincut_losses_only_backward(cuda, K, H):
    """The reference keeps only the two losses (hscn.py:63): the diag-only fast path (K < 64) and the split backward
    without pooled-feature / coarse-adjacency gradients (K >= 64)."""
    o, p = _oracle(), _product()
    b = _peptide_batch(9, seed=77)
    g = torch.Generator().manual_seed(1)
    N = b.x.size(0)
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    x = torch.randn(N, H, generator=g)
    s = torch.randn(N, K, generator=g)
    sr, st_ = s.clone().requires_grad_(), s.to(cuda).requires_grad_()
    _, _, mc_r, or_r = o.mincut_pool_ragged(x, ei, sr, b.batch)
    _, _, mc_t, or_t = p.mincut_pool_ragged(x.to(cuda), ei.to(cuda), st_, b.batch.to(cuda), want_out=False,
                                            want_adj=False)
    (mc_r + or_r).backward()
    (mc_t + or_t).backward()
    assert_close(mc_t, mc_r, RTOL, "mc")
    assert_close(or_t, or_r, RTOL, "ortho")
    assert_close(st_.grad, sr.grad, 5 * RTOL, "d logits")


@pytest.mark.parametrize("which", ["out", "adj"])
def test_mincut_split_backward_partial_objectives(cuda, which):
    """K >= 64: the split backward with only ONE of the two dense outputs in the objective (the other's gradient arrives
    as None) and features that carry no gradient (d_x not requested); rows of a trailing graph the batch vector does
    not cover get zero gradients."""
    o, p = _oracle(), _product()
    K, H = 64, 40
    b = _peptide_batch(5, seed=3)
    g = torch.Generator().manual_seed(7)
    N = b.x.size(0)
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    x = torch.randn(N, H, generator=g)
    s = torch.randn(N, K, generator=g)
    sr, st_ = s.clone().requires_grad_(), s.to(cuda).requires_grad_()
    out_r, adj_r, mc_r, or_r = o.mincut_pool_ragged(x, ei, sr, b.batch)
    out_t, adj_t, mc_t, or_t = p.mincut_pool_ragged(x.to(cuda), ei.to(cuda), st_, b.batch.to(cuda))
    w = torch.randn((out_r if which == "out" else adj_r).shape, generator=g)
    ((out_r if which == "out" else adj_r) * w).sum().add(mc_r * 0.5).backward()
    ((out_t if which == "out" else adj_t) * w.to(cuda)).sum().add(mc_t * 0.5).backward()
    assert_close(st_.grad, sr.grad, 5 * RTOL, f"d logits through {which}")


def test_dense_mincut_pool_single_graph_api(cuda):
    """Exactly the reference call pattern: to_dense_adj(edge_index) -> dense_mincut_pool(x, adj, s)."""
    o, p = _oracle(), _product()
    from graph_hscn_b200 import synthetic
    d = synthetic.peptides_graphs(1, seed=4)[0]
    n = d.num_nodes
    ei, _ = o.gcn_norm(d.edge_index, None, n, add_self_loops=True)
    g = torch.Generator().manual_seed(2)
    x, s = torch.randn(n, 16, generator=g), torch.randn(n, 5, generator=g)
    ref = o.dense_mincut_pool(x, o.to_dense_adj(ei), s)
    adj = p.to_dense_adj(ei.to(cuda))
    got = p.dense_mincut_pool(x.to(cuda), adj, s.to(cuda))
    for a, b_, nm in zip(got, ref, ["out", "out_adj", "mc", "ortho"]):
        assert_close(a, b_, RTOL, nm)
    assert got[0].shape == ref[0].shape and got[1].shape == ref[1].shape
    # the lazy adjacency densifies to the same tensor PyG would have returned
    assert torch.equal(adj.to_dense().cpu(), o.to_dense_adj(ei))
    # dense-input overload (a real dense tensor, weighted) + mask
    dense = torch.rand(2, 30, 30, generator=g) * (torch.rand(2, 30, 30, generator=g) < 0.2)
    xs, ss = torch.randn(2, 30, 8, generator=g), torch.randn(2, 30, 5, generator=g)
    mask = torch.ones(2, 30, dtype=torch.bool)
    mask[1, 20:] = False
    dense[1, 20:, :] = 0      # padded nodes of a collated batch carry no edges
    dense[1, :, 20:] = 0
    ref = o.dense_mincut_pool(xs, dense, ss, mask)
    got = p.dense_mincut_pool(xs.to(cuda), dense.to(cuda), ss.to(cuda), mask.to(cuda))
    for a, b_, nm in zip(got, ref, ["out", "out_adj", "mc", "ortho"]):
        assert_close(a, b_, RTOL, "dense " + nm)


def test_mincut_voc_sp_shape_vs_oracle(cuda):
    """BASELINE config #4 shape (n ~ 395-500, avg degree ~5.7, K=32, H=256): four sampled graphs, all four outputs
    and both gradients against the oracle's dense PyG recipe."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    o, p = _oracle(), _product()
    b = Batch.from_data_list(synthetic.vocsp_graphs(4, seed=1238))
    g = torch.Generator().manual_seed(4)
    N, K, H = b.x.size(0), 32, 256
    assert 4 * 395 <= N <= 4 * 500
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    x = torch.randn(N, H, generator=g)
    s = torch.randn(N, K, generator=g)
    xr, sr = x.clone().requires_grad_(), s.clone().requires_grad_()
    xt, st_ = x.to(cuda).requires_grad_(), s.to(cuda).requires_grad_()
    out_r, adj_r, mc_r, or_r = o.mincut_pool_ragged(xr, ei, sr, b.batch)
    out_t, adj_t, mc_t, or_t = p.mincut_pool_ragged(xt, ei.to(cuda), st_, b.batch.to(cuda))
    for a, c, nm in [(out_t, out_r, "out"), (adj_t, adj_r, "out_adj"), (mc_t, mc_r, "mc"), (or_t, or_r, "ortho")]:
        assert_close(a, c, RTOL, "VOC-SP mincut " + nm)
    go = torch.randn(out_r.shape, generator=g)
    (mc_r + or_r + (out_r * go).sum() * 0.01).backward()
    (mc_t + or_t + (out_t * go.to(cuda)).sum() * 0.01).backward()
    assert_close(st_.grad, sr.grad, 5 * RTOL, "VOC-SP mincut d logits")
    assert_close(xt.grad, xr.grad, 5 * RTOL, "VOC-SP mincut d x")


@pytest.mark.parametrize("train_eps", [False, True])
def test_ginconv_fwd_bwd(cuda, train_eps):
    """GINConv is listed in CONV_DICT (config/config.py:19-23): out = nn((1 + eps) x_i + sum_j x_j)."""
    import torch.nn as nn
    o, p = _oracle(), _product()
    g = torch.Generator().manual_seed(12)
    n, e, fi, fo = 700, 2600, 9, 24
    ei = random_edge_index(n, n, e, g)
    x = torch.randn(n, fi, generator=g)

    def mlp(ns):
        return nn.Sequential(ns.Linear(fi, 32), nn.ReLU(), ns.Linear(32, fo))
    ref, tst = _to_dev(lambda: o.GINConv(mlp(o), eps=0.3, train_eps=train_eps),
                       lambda: p.GINConv(mlp(p), eps=0.3, train_eps=train_eps), cuda, lambda m, d: None)
    xr, xt = x.clone().requires_grad_(), x.to(cuda).requires_grad_()
    yr, yt = ref(xr, ei), tst(xt, ei.to(cuda))
    assert_close(yt, yr, RTOL, "GIN out")
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xt.grad, xr.grad, 2 * RTOL, "GIN dx")
    for (nm, a), (_, c) in zip(tst.named_parameters(), ref.named_parameters()):
        assert_close(a.grad, c.grad, 2 * RTOL, f"GIN d{nm}")


def test_adamw_kernel_matches_torch_adamw(cuda):
    """ghscn_adamw_step (flat buffer, device step counter) vs torch.optim.AdamW over 5 steps at 1e-6."""
    from graph_hscn_b200.train import FlatAdamW
    g = torch.Generator().manual_seed(13)
    n = 277_001
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)
    p = p0.to(cuda)
    grad = torch.empty(n, device=cuda)
    mine = FlatAdamW(p, grad, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)
    for step in range(5):
        gr = torch.randn(n, generator=g) * (10.0 ** (step - 2))
        ref.grad = gr.clone()
        opt.step()
        grad.copy_(gr)
        mine.step()
    assert_close(p, ref.detach(), 1e-6, "parameters after 5 AdamW steps")
    st = opt.state[ref]
    assert_close(mine.exp_avg, st["exp_avg"], 1e-6, "exp_avg")
    assert_close(mine.exp_avg_sq, st["exp_avg_sq"], 1e-6, "exp_avg_sq")
    assert float(mine.step_count) == 5.0


# ------------------------------------------------------------------------------------------- K7
@pytest.mark.parametrize("K", [4, 10, 32])
@pytest.mark.parametrize("int_feats", [True, False])
def test_virtual_nodes_bit_exact(cuda, K, int_feats):
    from graph_hscn_b200 import hetero, synthetic
    from graph_hscn_b200.data import Batch, HeteroData
    from oracle import hetero as ohet
    graphs = synthetic.peptides_graphs(7, seed=K)
    g = torch.Generator().manual_seed(K)
    if not int_feats:
        for d in graphs:
            d.x = torch.randn(d.x.shape, generator=g)
    b = Batch.from_data_list(graphs)
    N = b.x.size(0)
    s = torch.softmax(torch.randn(N, K, generator=g) * 3, -1)
    s[3] = s[3, 0]          # an exact tie row: first index must win
    clusters_ref = ohet.cluster_argmax(s)
    clusters = hetero.assign_clusters(s.to(cuda))
    assert np.array_equal(clusters.cpu().numpy(), clusters_ref)
    if K == 4:
        clusters_ref[: graphs[0].num_nodes] = 2      # a graph with a single non-empty cluster
        clusters = torch.from_numpy(clusters_ref).int().to(cuda)
    # reference: per-graph HeteroData + PyG collate
    hlist = []
    for gi, d in enumerate(graphs):
        lo, hi = int(b.ptr[gi]), int(b.ptr[gi + 1])
        _, vx, vv, lv = ohet.virtual_nodes(d.x, clusters_ref[lo:hi], K)
        h = HeteroData()
        h["local"].x = d.x.float()
        h["local"].y = d.y
        h["virtual"].x = vx
        h["local", "to", "local"].edge_index = d.edge_index
        h["virtual", "to", "virtual"].edge_index = vv
        h["local", "to", "virtual"].edge_index = lv
        hlist.append(h)
    ref = Batch.from_data_list(hlist)
    got = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters, K,
                                    y=b.y.to(cuda))
    assert list(got.edge_index_dict.keys()) == list(ref.edge_index_dict.keys())
    for et in ref.edge_index_dict:
        assert torch.equal(got.edge_index_dict[et].cpu(), ref.edge_index_dict[et]), et
    assert torch.equal(got["virtual"].x.cpu(), ref["virtual"].x), "virtual features must be bit-exact"
    assert torch.equal(got["virtual"].batch.cpu(), ref["virtual"].batch)
    assert torch.equal(got["local"].x.cpu(), ref["local"].x)
    # padded layout == compact layout after dropping the empty slots
    pad = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters, K,
                                    padded=True)
    nv = got.num_virtual.cpu()
    sel = torch.cat([gi * K + torch.arange(int(u)) for gi, u in enumerate(nv)])
    assert torch.equal(pad["virtual"].x.cpu()[sel], ref["virtual"].x)
    vvp = pad[("virtual", "to", "virtual")].edge_index.cpu()
    assert int((vvp[0] >= 0).sum()) == ref[("virtual", "to", "virtual")].edge_index.size(1)


# ------------------------------------------------------------------------------------- model level
def test_mpnn_model_fwd_bwd(cuda):
    """Config #1 shape at reduced width: MPNN(GCN) on a Peptides-shaped batch, BCE loss."""
    from graph_hscn_b200 import models
    o, p = _oracle(), _product()
    b = _peptide_batch(16, seed=21)
    b.x = b.x.float()
    ref, tst = _to_dev(lambda: models.MPNN("gcn", torch.relu, 9, 64, 10, 5, ops=o),
                       lambda: models.MPNN("gcn", torch.relu, 9, 64, 10, 5, ops=p), cuda, lambda m, d: None)
    bt = b.to(cuda)
    yr, yt = ref(b), tst(bt)
    assert_close(yt, yr, RTOL, "MPNN logits")
    lr, _ = models.criterion("cross_entropy", yr, b.y)
    lt, _ = models.criterion("cross_entropy", yt, bt.y)
    lr.backward()
    lt.backward()
    assert_close(lt, lr, RTOL, "loss")
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), tst.named_parameters()):
        assert_close(p2.grad, p1.grad, 10 * RTOL, f"grad {n1}")


def test_scn_per_graph_and_batched(cuda):
    from graph_hscn_b200 import models, synthetic
    o, p = _oracle(), _product()
    ref, tst = _to_dev(lambda: models.SCN([16], "elu", 9, 10, ops=o), lambda: models.SCN([16], "elu", 9, 10, ops=p),
                       cuda, lambda m, d: None)
    d = synthetic.peptides_graphs(1, seed=31)[0]
    n = d.num_nodes
    ei_r, w_r = o.gcn_norm(d.edge_index, None, n, add_self_loops=True)
    ei_t, w_t = p.gcn_norm(d.edge_index.to(cuda), None, n, add_self_loops=True)
    Sr, mcr, orr, adjr = ref(d.x.float(), ei_r, w_r)
    St, mct, ort, adjt = tst(d.x.float().to(cuda), ei_t, w_t)
    assert_close(St, Sr, RTOL, "S")
    assert_close(mct, mcr, RTOL, "mc")
    assert_close(ort, orr, RTOL, "ortho")
    (mcr + orr).backward()
    (mct + ort).backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), tst.named_parameters()):
        assert_close(p2.grad, p1.grad, 20 * RTOL, f"SCN grad {n1}")
    # cluster ids: bit-exact wherever the fp32 softmax margin exceeds rounding noise
    top2 = Sr.topk(2, dim=1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 1e-5
    from graph_hscn_b200 import hetero
    cl = hetero.assign_clusters(St).cpu()
    assert torch.equal(cl[safe].long(), Sr.max(1)[1][safe])
    assert float(safe.float().mean()) > 0.9
    # batched
    b = _peptide_batch(12, seed=32)
    N = b.x.size(0)
    ref.zero_grad(); tst.zero_grad()
    ei_r, w_r = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    ei_t, w_t = p.gcn_norm(b.edge_index.to(cuda), None, N, add_self_loops=True)
    _, mcr, orr = ref.forward_batched(b.x.float(), ei_r, w_r, b.batch)
    _, mct, ort = tst.forward_batched(b.x.float().to(cuda), ei_t, w_t, b.batch.to(cuda))
    assert_close(mct, mcr, RTOL, "batched mc")
    assert_close(ort, orr, RTOL, "batched ortho")
    (mcr + orr).backward()
    (mct + ort).backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), tst.named_parameters()):
        assert_close(p2.grad, p1.grad, 20 * RTOL, f"SCN batched grad {n1}")


@pytest.mark.parametrize("act,units,K", [("elu", 16, 10), ("relu", 32, 32), ("tanh", 8, 5), ("identity", 16, 10)])
def test_scn_fused_node_pipeline(cuda, act, units, K):
    """ghscn_scn_forward (GraphConv aggregation + lin_rel + lin_root + activation + cluster Linear in one launch)
    against the separate operators of this package and against the CPU oracle, values and parameter gradients."""
    from graph_hscn_b200 import models
    o, p = _oracle(), _product()
    ref, tst = _to_dev(lambda: models.SCN([units], act, 9, K, ops=o), lambda: models.SCN([units], act, 9, K, ops=p),
                       cuda, lambda m, d: None)
    assert tst._fusable and not ref._fusable
    b = _peptide_batch(12, seed=33)
    N = b.x.size(0)
    g = torch.Generator().manual_seed(units)
    up = torch.randn(N, K, generator=g)
    ei_r, w_r = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    ei_t, w_t = p.gcn_norm(b.edge_index.to(cuda), None, N, add_self_loops=True)
    x_t = b.x.float().to(cuda)
    s_r = ref.logits(b.x.float(), ei_r, w_r)
    (s_r * up).sum().backward()
    got = {}
    from graph_hscn_b200 import ops as gops
    for fuse in (True, False, "chain"):                    # fused fwd + fused bwd | separate operators | fused fwd + op chain bwd
        tst.zero_grad(set_to_none=True)
        tst.fuse = bool(fuse)
        gops.FUSED_SCN_BACKWARD = fuse is True
        try:
            s_t = tst.logits(x_t, ei_t, w_t)
            (s_t * up.to(cuda)).sum().backward()
        finally:
            gops.FUSED_SCN_BACKWARD = True
        got[fuse] = (s_t.detach().clone(), {n: q.grad.detach().clone() for n, q in tst.named_parameters()})
    for n in got[True][1]:
        assert_close(got[True][1][n], got["chain"][1][n], RTOL, f"fused backward vs operator chain, grad {n}")
    assert_close(got[True][0], got[False][0], 2e-6, "fused vs separate logits")
    assert_close(got[True][0], s_r, RTOL, "fused logits vs oracle")
    for n, q in ref.named_parameters():
        assert_close(got[True][1][n], got[False][1][n], RTOL, f"fused vs separate grad {n}")
        assert_close(got[True][1][n], q.grad, 10 * RTOL, f"fused grad {n} vs oracle")
    with torch.no_grad():                                  # the assignment pass: nothing saved, same logits
        tst.fuse = True
        assert torch.equal(tst.logits(x_t, ei_t, w_t), got[True][0])


def test_hscn_model_fwd_bwd(cuda):
    """Config #2 shape at reduced width: 3-relation HeteroConv (GAT l->v, GCN l->l, GCN v->v)."""
    from graph_hscn_b200 import hetero, models
    o, p = _oracle(), _product()
    K = 10
    b = _peptide_batch(10, seed=41)
    g = torch.Generator().manual_seed(41)
    clusters = torch.randint(0, K, (b.x.size(0),), generator=g).int()
    hb = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters.to(cuda), K,
                                   y=b.y.to(cuda))
    hb_cpu = hb.to("cpu")

    def prime(m, d):
        src = hb_cpu if d == "cpu" else hb
        m(src.x_dict, src.edge_index_dict, src)
    ref, tst = _to_dev(lambda: models.HSCN("GAT", "GCN", "GCN", torch.relu, 9, 48, 10, 3, ops=o),
                       lambda: models.HSCN("GAT", "GCN", "GCN", torch.relu, 9, 48, 10, 3, ops=p), cuda, prime)
    yr = ref(hb_cpu.x_dict, hb_cpu.edge_index_dict, hb_cpu)
    yt = tst(hb.x_dict, hb.edge_index_dict, hb)
    assert_close(yt, yr, RTOL, "HSCN logits")
    # operator-level parity of the virtual branch (dead w.r.t. the loss but part of HeteroConv's output)
    xr = ref.convs[0](hb_cpu.x_dict, hb_cpu.edge_index_dict)
    xt = tst.convs[0](hb.x_dict, hb.edge_index_dict)
    assert set(xr) == set(xt) == {"local", "virtual"}
    want_virtual = xr["virtual"].relu() if tst._fused_virtual else xr["virtual"]
    assert_close(xt["virtual"], want_virtual, RTOL, "HeteroConv virtual")
    # the mirror model lets the l->l GCN apply the following ReLU in its aggregation epilogue
    want_local = xr["local"].relu() if tst._fused_local else xr["local"]
    assert_close(xt["local"], want_local, RTOL, "HeteroConv local")
    lr, _ = models.criterion("cross_entropy", yr, hb_cpu["local"].y)
    lt, _ = models.criterion("cross_entropy", yt, hb["local"].y)
    lr.backward()
    lt.backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), tst.named_parameters()):
        if p1.grad is None:
            assert p2.grad is None or float(p2.grad.abs().max()) == 0.0, n1   # dead virtual branch
        else:
            assert_close(p2.grad, p1.grad, 10 * RTOL, f"HSCN grad {n1}")


@pytest.mark.parametrize("width,fuse_relu", [(9, False), (48, True), (300, True)])
def test_fused_virtual_layer_matches_unfused_and_oracle(cuda, width, fuse_relu):
    """HeteroConv computes the "virtual" destination (v->v GCN + l->v GAT pool, summed) with ONE fused operator at the
    input width: outputs vs the oracle's HeteroConv, and -- through the fused operator's backward -- the gradients of
    every virtual-branch parameter and of both inputs."""
    from graph_hscn_b200 import hetero
    from graph_hscn_b200.pyg import nn as pnn
    o, p = _oracle(), _product()
    K, H = 10, 64
    b = _peptide_batch(9, seed=17)
    g = torch.Generator().manual_seed(width)
    clusters = torch.randint(0, K, (b.x.size(0),), generator=g).int()
    hb = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters.to(cuda), K)
    hb_cpu = hb.to("cpu")
    N, V = hb_cpu["local"].x.size(0), hb_cpu["virtual"].x.size(0)
    xl, xv = torch.randn(N, width, generator=g), torch.randn(V, width, generator=g)

    def build(ns):
        return ns.HeteroConv({("local", "to", "virtual"): ns.GATConv((-1, -1), H, add_self_loops=False),
                              ("local", "to", "local"): ns.GCNConv(-1, H, add_self_loops=False),
                              ("virtual", "to", "virtual"): ns.GCNConv(-1, H, add_self_loops=False)}, aggr="sum")
    ref, tst = _to_dev(lambda: build(o), lambda: build(p), cuda,
                       lambda m, d: m({"local": xl.to(d), "virtual": xv.to(d)},
                                      hb_cpu.edge_index_dict if d == "cpu" else hb.edge_index_dict))
    if fuse_relu:
        tst.fuse_relu_dst = {"virtual"}
    xlr, xvr = xl.clone().requires_grad_(), xv.clone().requires_grad_()
    xlt, xvt = xl.to(cuda).requires_grad_(), xv.to(cuda).requires_grad_()
    assert tst._fused_virtual_plan({"local": xlt, "virtual": xvt}, hb.edge_index_dict), "fused path not taken"
    yr = ref({"local": xlr, "virtual": xvr}, hb_cpu.edge_index_dict)["virtual"]
    if fuse_relu:
        yr = yr.relu()
    yt = tst({"local": xlt, "virtual": xvt}, hb.edge_index_dict)["virtual"]
    assert_close(yt, yr, RTOL, "fused virtual output")
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    yt.backward(gy.to(cuda))
    assert_close(xlt.grad, xlr.grad, 5 * RTOL, "d x_local")
    assert_close(xvt.grad, xvr.grad, 5 * RTOL, "d x_virtual")
    for (n1, p1), (_, p2) in zip(ref.named_parameters(), tst.named_parameters()):
        if "local__to__local" in n1:
            continue
        assert_close(p2.grad, p1.grad, 5 * RTOL, f"virtual-branch grad {n1}")
    # and the unfused schedule gives the same numbers
    old = pnn.FUSED_VIRTUAL
    pnn.FUSED_VIRTUAL = False
    try:
        yu = tst({"local": xlt.detach(), "virtual": xvt.detach()}, hb.edge_index_dict)["virtual"]
    finally:
        pnn.FUSED_VIRTUAL = old
    assert_close(yu, yt.detach(), RTOL, "fused vs unfused")


def test_cpu_tensor_fails_loudly(cuda):
    from graph_hscn_b200 import pyg
    conv = pyg.GCNConv(4, 4)
    with pytest.raises(RuntimeError):
        conv(torch.randn(3, 4), torch.tensor([[0, 1], [1, 2]]))
