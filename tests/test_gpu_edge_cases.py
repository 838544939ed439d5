"""GPU edge cases: empty / ragged inputs, isolated nodes, tiny and maximum-size graphs, existing self loops."""
import pytest
import torch

from tests.util import RTOL, assert_close

pytestmark = pytest.mark.gpu


def _ns():
    from graph_hscn_b200 import pyg
    from oracle.namespace import namespace
    return namespace(), pyg.namespace()


def _pair(ctor_o, ctor_p, dev):
    torch.manual_seed(0)
    ref = ctor_o()
    tst = ctor_p().to(dev)
    tst.load_state_dict(ref.state_dict())
    return ref, tst


def test_gcnconv_without_edges_and_with_isolated_nodes(cuda):
    o, p = _ns()
    ref, tst = _pair(lambda: o.GCNConv(5, 8), lambda: p.GCNConv(5, 8), cuda)
    x = torch.randn(6, 5)
    empty = torch.zeros(2, 0, dtype=torch.long)
    assert_close(tst(x.to(cuda), empty.to(cuda)), ref(x, empty), RTOL, "E=0 with self loops")
    ref2, tst2 = _pair(lambda: o.GCNConv(5, 8, add_self_loops=False), lambda: p.GCNConv(5, 8, add_self_loops=False), cuda)
    ei = torch.tensor([[0, 1], [1, 0]])                    # nodes 2..5 isolated: deg 0 -> inf -> 0 rule
    assert_close(tst2(x.to(cuda), ei.to(cuda)), ref2(x, ei), RTOL, "isolated nodes, no self loops")
    assert_close(tst2(x.to(cuda), empty.to(cuda)), ref2(x, empty), RTOL, "E=0, no self loops")


def test_gcnconv_with_existing_self_loops_and_duplicates(cuda):
    o, p = _ns()
    ref, tst = _pair(lambda: o.GCNConv(4, 6), lambda: p.GCNConv(4, 6), cuda)
    ei = torch.tensor([[0, 1, 1, 2, 2, 2, 3, 0], [1, 0, 1, 2, 0, 0, 3, 1]])   # loops at 1,2,3 + duplicate (2,0),(0,1)
    w = torch.tensor([0.5, 2.0, 3.0, 4.0, 1.5, 0.25, 7.0, 1.0])
    x = torch.randn(4, 4)
    assert_close(tst(x.to(cuda), ei.to(cuda)), ref(x, ei), RTOL, "unweighted")
    assert_close(tst(x.to(cuda), ei.to(cuda), w.to(cuda)), ref(x, ei, w), RTOL, "weighted, existing loops keep their weight")


def test_ragged_batch_with_tiny_and_max_size_graphs(cuda):
    """1-node graph without edges, 2-node graph, and a 444-node graph (the Peptides maximum) in one batch."""
    from graph_hscn_b200 import hetero, models, synthetic
    from graph_hscn_b200.data import Batch, Data
    from oracle import hetero as ohet
    o, p = _ns()
    g = torch.Generator().manual_seed(1)
    one = Data(x=torch.randint(0, 5, (1, 9), generator=g), edge_index=torch.zeros(2, 0, dtype=torch.long),
               y=torch.zeros(1, 10))
    two = Data(x=torch.randint(0, 5, (2, 9), generator=g), edge_index=torch.tensor([[0, 1], [1, 0]]), y=torch.ones(1, 10))
    big = synthetic.peptides_graphs(1, seed=3, fixed_nodes=444)[0]
    mid = synthetic.peptides_graphs(1, seed=4)[0]
    b = Batch.from_data_list([one, big, two, mid])
    bx = b.x.float()
    # MPNN
    ref, tst = _pair(lambda: models.MPNN("gcn", torch.relu, 9, 32, 10, 3, ops=o),
                     lambda: models.MPNN("gcn", torch.relu, 9, 32, 10, 3, ops=p), cuda)
    b.x = bx
    assert_close(tst(b.to(cuda)), ref(b), RTOL, "MPNN on ragged batch")
    # MinCUT with K larger than the smallest graphs
    K = 6
    N = bx.size(0)
    ei, _ = o.gcn_norm(b.edge_index, None, N, add_self_loops=True)
    s = torch.randn(N, K, generator=g)
    xr = torch.randn(N, 16, generator=g)
    out_r = o.mincut_pool_ragged(xr, ei, s, b.batch)
    out_t = p.mincut_pool_ragged(xr.to(cuda), ei.to(cuda), s.to(cuda), b.batch.to(cuda))
    for a, c, nm in zip(out_t, out_r, ["out", "out_adj", "mc", "ortho"]):
        assert_close(a, c, RTOL, f"ragged mincut {nm}")
    # virtual nodes: graphs with fewer nodes than clusters
    clusters = torch.softmax(s * 4, -1).max(1)[1]
    x_raw = torch.cat([d.x for d in [one, big, two, mid]])             # int64 atom features
    hb = hetero.build_hetero_batch(x_raw.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda),
                                   clusters.int().to(cuda), K)
    off = voff = 0
    for d in [one, big, two, mid]:
        n = d.num_nodes
        _, vx, vv, lv = ohet.virtual_nodes(d.x, clusters[off:off + n].numpy(), K)
        U = vx.size(0)
        assert torch.equal(hb["virtual"].x[voff:voff + U].cpu(), vx)
        off, voff = off + n, voff + U
    assert voff == hb["virtual"].x.size(0)


def test_pool_with_empty_graphs_in_the_middle(cuda):
    o, p = _ns()
    x = torch.randn(7, 12)
    batch = torch.tensor([0, 0, 2, 2, 2, 5, 5])            # graphs 1, 3, 4 have no nodes
    assert_close(p.global_mean_pool(x.to(cuda), batch.to(cuda), 7), o.global_mean_pool(x, batch, 7), 1e-6, "empty segments")
    assert_close(p.scatter_mean(x.to(cuda), batch.to(cuda), dim=0), o.scatter_mean(x, batch, dim=0), 1e-6, "inferred size")


def test_gat_destination_without_members_gets_bias_only(cuda):
    o, p = _ns()
    torch.manual_seed(0)
    ref = o.GATConv((4, 4), 8, add_self_loops=False)
    tst = p.GATConv((4, 4), 8, add_self_loops=False).to(cuda)
    tst.load_state_dict(ref.state_dict())
    with torch.no_grad():
        ref.bias.uniform_(-1, 1)
        tst.bias.copy_(ref.bias)
    xs, xd = torch.randn(5, 4), torch.randn(3, 4)
    ei = torch.tensor([[0, 1, 2, 3, 4], [0, 0, 2, 2, 2]])   # destination 1 has no incoming edge
    yr, yt = ref((xs, xd), ei), tst((xs.to(cuda), xd.to(cuda)), ei.to(cuda))
    assert_close(yt, yr, RTOL, "GAT with an empty destination")
    assert torch.allclose(yt[1].cpu(), ref.bias)



def test_mincut_stale_max_nodes_hint_poisons_instead_of_overrunning(cuda):
    """A max-nodes-per-graph hint smaller than the largest graph (the hint sizes the kernels' shared tiles and grids):
    the affected graph's losses, pooled features and gradients come back NaN -- nothing is written past a tile and the
    graphs inside the hint are untouched."""
    from graph_hscn_b200 import pyg, synthetic
    from graph_hscn_b200.structure import structure_cache, structure_hints
    b = synthetic.peptides_batch(4, seed=11)
    counts = (b.ptr[1:] - b.ptr[:-1])
    big = int(counts.argmax())
    small_cap = int(counts.sort().values[-2])            # fits every graph but the largest
    assert small_cap < int(counts.max())
    N, K, H = b.x.size(0), 10, 128
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, H, generator=g).to(cuda).requires_grad_()
    s = torch.randn(N, K, generator=g).to(cuda).requires_grad_()
    ei, _ = pyg.gcn_norm(b.edge_index.to(cuda), None, N, add_self_loops=True)     # structures built without the hint
    with structure_hints(num_graphs=4, batch_sorted=1, max_nodes_per_graph=small_cap):
        out, adj, mc, ol = pyg.mincut_pool_ragged(x, ei, s, b.batch.to(cuda))
        (mc + ol + out.sum() + adj.sum()).backward()
    torch.cuda.synchronize()
    assert bool(torch.isnan(mc)) and bool(torch.isnan(ol))
    assert bool(torch.isnan(out[big]).all())
    ok = [i for i in range(4) if i != big]
    assert bool(torch.isfinite(out[ok]).all())
    lo, hi = int(b.ptr[big]), int(b.ptr[big + 1])
    assert bool(torch.isnan(s.grad[lo:hi]).any()) and bool(torch.isnan(x.grad[lo:hi]).any())
    structure_cache().clear()
