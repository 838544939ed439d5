"""Full-size GPU tests (BASELINE.json configs #3 and #4): the CPU oracle would need minutes at these sizes, so the
kernels are checked through size-independent properties of the domain -- sortedness / permutation validity of the
CSR, linearity and row sums of the aggregation, invariance of the batch-mean MinCUT losses under a permutation of
the graphs, conservation of feature mass by the virtual-node means, consistency of sharded and full-batch results
(the data-parallel identity of SURVEY 8e) -- plus a sampled exact comparison against the oracle on a few graphs cut
out of the big batch."""
import pytest
import torch

from tests.util import RTOL, assert_close

pytestmark = pytest.mark.gpu


def _ops():
    from graph_hscn_b200 import pyg
    return pyg.namespace()


@pytest.fixture(scope="module")
def struct_batch():
    """Config #3: Peptides-struct shape, 1024 graphs (~155 k nodes, ~316 k directed edges)."""
    from graph_hscn_b200 import synthetic
    return synthetic.peptides_batch(1024, seed=1234 + 3, task="struct")


@pytest.fixture(scope="module")
def voc_batch():
    """Config #4: PascalVOC-SP shape, 128 graphs of 395..500 nodes, average degree ~5.7, 14 features."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    return Batch.from_data_list(synthetic.vocsp_graphs(128, seed=1238))


@pytest.mark.parametrize("which", ["struct", "voc"])
def test_csr_full_size_is_a_stable_sort(cuda, which, struct_batch, voc_batch):
    """rowptr monotone and complete, perm a permutation, keys sorted, ties in edge order; the per-graph fast path
    agrees with the radix path at full size."""
    from graph_hscn_b200.structure import StructureCache, build_csr, edge_blocks_from_batch
    b = struct_batch if which == "struct" else voc_batch
    ei, N, E = b.edge_index.to(cuda), b.x.size(0), b.edge_index.size(1)
    for key, other in ((ei[1], ei[0]), (ei[0], ei[1])):
        c = build_csr(key, other, N, False)
        rp, perm, col = c.rowptr.long(), c.perm.long(), c.col.long()
        assert int(rp[0]) == 0 and int(rp[-1]) == E and bool((rp[1:] >= rp[:-1]).all())
        assert torch.equal(torch.sort(perm)[0], torch.arange(E, device=cuda))
        ks = key[perm]
        assert bool((ks[1:] >= ks[:-1]).all())                                     # sorted by key
        assert bool(((ks[1:] > ks[:-1]) | (perm[1:] > perm[:-1])).all())           # stable inside a row
        assert torch.equal(col, other[perm])
        assert torch.equal(torch.bincount(key, minlength=N), rp[1:] - rp[:-1])     # row lengths = in/out degrees
    blocks = edge_blocks_from_batch(b.edge_index, b.batch, int(b.num_graphs))
    assert blocks is not None
    cache = StructureCache()
    batch_d = b.batch.to(cuda)
    cache.blocked_status(cuda).zero_()
    cache.register_blocks(ei, cache.segments(batch_d, int(b.num_graphs)).ptr, int(b.num_graphs), *blocks)
    st = cache.graph(ei, N, N, False)
    for fast, (key, other) in ((st.by_dst, (ei[1], ei[0])), (st.by_src, (ei[0], ei[1]))):
        ref = build_csr(key, other, N, False)
        assert torch.equal(fast.rowptr, ref.rowptr) and torch.equal(fast.col, ref.col) and torch.equal(fast.perm, ref.perm)
    assert int(cache.blocked_status(cuda)) == 0


@pytest.mark.parametrize("which,width", [("struct", 300), ("voc", 256)])
def test_gcn_aggregation_full_size_linearity_and_row_sums(cuda, which, width, struct_batch, voc_batch):
    """GCNConv without its projection is the linear map D^-1/2 (A + I) D^-1/2: additive, homogeneous, and on the
    all-ones vector it returns the weighted degree, which gcn_norm's own edge weights give independently."""
    p = _ops()
    b = struct_batch if which == "struct" else voc_batch
    N = b.x.size(0)
    ei = b.edge_index.to(cuda)
    g = torch.Generator(device="cuda").manual_seed(5)
    conv = p.GCNConv(width, width, bias=False).to(cuda)
    with torch.no_grad():
        conv.lin.weight.copy_(torch.eye(width, device=cuda))
        x = torch.randn(N, width, device=cuda, generator=g)
        y = torch.randn(N, width, device=cuda, generator=g)
        fx, fy, fxy = conv(x, ei), conv(y, ei), conv(1.5 * x - 0.25 * y, ei)
        assert_close(fxy, 1.5 * fx - 0.25 * fy, 5 * RTOL, "linearity of the aggregation")
        ones = conv(torch.ones(N, width, device=cuda), ei)
        ei2, w = p.gcn_norm(ei, None, N, add_self_loops=True)
        deg = torch.zeros(N, device=cuda, dtype=torch.float64).index_add_(0, ei2[1], w.double())
        assert_close(ones[:, 0], deg.float(), RTOL, "row sums = weighted in-degree")
        assert_close(ones[:, -1], deg.float(), RTOL, "row sums (last column)")


def test_mincut_full_size_ranges_permutation_and_sampled_oracle(cuda, struct_batch):
    """Batch-mean MinCUT / orthogonality losses at B = 1024: inside their analytic ranges, unchanged when the graphs
    are fed in another order, equal to the mean of two half-batch calls, and the per-graph pooled outputs of a few
    sampled graphs equal the CPU oracle on the same graph alone."""
    from graph_hscn_b200.data import Batch
    from oracle.namespace import namespace as oracle_ns
    p, o = _ops(), oracle_ns()
    K, H = 10, 300
    graphs = struct_batch.to_data_list()
    g = torch.Generator().manual_seed(9)
    logits = [torch.randn(d.num_nodes, K, generator=g) for d in graphs]
    feats = [torch.randn(d.num_nodes, H, generator=g) for d in graphs]

    def run(order):
        b = Batch.from_data_list([graphs[i] for i in order])
        N = b.x.size(0)
        ei, _ = p.gcn_norm(b.edge_index.to(cuda), None, N, add_self_loops=True)
        s = torch.cat([logits[i] for i in order]).to(cuda)
        x = torch.cat([feats[i] for i in order]).to(cuda)
        return p.mincut_pool_ragged(x, ei, s, b.batch.to(cuda))

    order = list(range(len(graphs)))
    out, adj, mc, orth = run(order)
    assert -1.0 - 1e-6 <= float(mc) <= 0.0 and 0.0 <= float(orth) <= 2.0
    perm = torch.randperm(len(graphs), generator=g).tolist()
    out_p, adj_p, mc_p, orth_p = run(perm)
    assert_close(mc_p, mc, RTOL, "mincut loss under a permutation of the graphs")
    assert_close(orth_p, orth, RTOL, "ortho loss under a permutation of the graphs")
    assert_close(out_p, out[perm], 1e-6, "per-graph pooled features do not depend on the graph's position")
    assert_close(adj_p, adj[perm], 1e-6, "per-graph pooled adjacency does not depend on the graph's position")
    half = len(graphs) // 2
    _, _, mc_a, or_a = run(order[:half])
    _, _, mc_b, or_b = run(order[half:])
    assert_close((mc_a + mc_b) / 2, mc, RTOL, "mean of the two shards' losses (data-parallel identity)")
    assert_close((or_a + or_b) / 2, orth, RTOL, "mean of the two shards' ortho losses")
    for i in (0, 17, 511, 1023):
        d = graphs[i]
        ei_r, _ = o.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
        out_r, adj_r, _, _ = o.mincut_pool_ragged(feats[i], ei_r, logits[i], torch.zeros(d.num_nodes, dtype=torch.long))
        assert_close(out[i:i + 1], out_r, RTOL, f"pooled features of graph {i}")
        assert_close(adj[i:i + 1], adj_r, RTOL, f"pooled adjacency of graph {i}")


def test_virtual_nodes_full_size_conserve_feature_mass(cuda, struct_batch):
    """Cluster means times cluster sizes add up to each graph's column sums (integer features: exactly), every local
    node has exactly one l->v edge, and the v->v pattern has U(U+1)/2 edges per graph."""
    from graph_hscn_b200 import hetero
    b = struct_batch
    K, B = 10, int(b.num_graphs)
    g = torch.Generator().manual_seed(4)
    clusters = torch.randint(0, K, (b.x.size(0),), generator=g).int()
    hb = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters.to(cuda), K)
    lv = hb["local", "to", "virtual"].edge_index
    vv = hb["virtual", "to", "virtual"].edge_index
    vx = hb["virtual"].x.double()
    vbatch = hb["virtual"].batch
    N = b.x.size(0)
    assert lv.size(1) == N and torch.equal(lv[0], torch.arange(N, device=cuda))
    sizes = torch.bincount(lv[1], minlength=vx.size(0)).double()
    assert bool((sizes > 0).all())                                                  # empty clusters are dropped
    # hetero_data.py:52-59 buckets cluster c into slot c - 1: virtual row j carries the mean of cluster (j + 1) mod U
    vptr = hb["virtual"].ptr.long()
    U = vptr[1:] - vptr[:-1]
    j = torch.arange(vx.size(0), device=cuda) - vptr[vbatch]
    sizes = sizes[vptr[vbatch] + (j + 1) % U[vbatch]]
    mass = torch.zeros(B, vx.size(1), device=cuda, dtype=torch.float64).index_add_(0, vbatch, vx * sizes[:, None])
    want = torch.zeros(B, vx.size(1), device=cuda, dtype=torch.float64).index_add_(0, b.batch.to(cuda), b.x.to(cuda).double())
    assert float((mass - want).abs().max()) <= 1e-3 * float(want.abs().max())       # fp32 means: rounding only
    assert vv.size(1) == int((U * (U + 1) // 2).sum())
    assert torch.equal(vbatch[lv[1]], b.batch.to(cuda))                             # membership stays inside the graph


def test_step_full_size_shards_reproduce_the_full_batch_gradient(cuda, struct_batch):
    """graphs are independent, so the HSCN gradient of the full batch equals the graph-count-weighted mean
    of the gradients of its shards (what the NCCL all-reduce computes), here on one GPU at B = 1024."""
    from graph_hscn_b200 import hetero, models
    from graph_hscn_b200.data import Batch
    p = _ops()
    K = 10
    graphs = struct_batch.to_data_list()
    torch.manual_seed(0)
    model = models.HSCN("GAT", "GCN", "GCN", torch.relu, 9, 64, 11, 2, ops=p).to(cuda)

    def grads(sub):
        b = Batch.from_data_list(sub).to(cuda)
        clusters = (torch.arange(b.x.size(0)) % K).int().to(cuda)
        hb = hetero.build_hetero_batch(b.x, b.edge_index, b.batch, clusters, K, y=b.y)
        model.zero_grad(set_to_none=True)
        pred = model(hb.x_dict, hb.edge_index_dict, hb)
        loss, _ = models.criterion("l1", pred, hb["local"].y)
        loss.backward()
        return {n: q.grad.detach().clone() for n, q in model.named_parameters() if q.grad is not None}, float(loss.detach())

    full, loss_full = grads(graphs)
    a, loss_a = grads(graphs[:384])
    c, loss_c = grads(graphs[384:])
    wa, wc = 384 / 1024, 640 / 1024
    assert abs(wa * loss_a + wc * loss_c - loss_full) <= 1e-5 * abs(loss_full)
    assert set(full) == set(a) == set(c) and len(full) > 0
    for n in full:
        assert_close(wa * a[n] + wc * c[n], full[n], 2e-4, f"sharded gradient of {n}")
