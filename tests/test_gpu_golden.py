"""GPU tests against the committed golden vectors (outputs of the UNMODIFIED reference sources running on
the CPU oracle, tests/golden/make_golden.py).  fp32 tolerance 1e-5 relative; integers bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.test_golden_cpu import graphs_from, load
from tests.util import RTOL, assert_close

pytestmark = pytest.mark.gpu


def _product():
    from graph_hscn_b200 import pyg
    return pyg.namespace()


def test_mpnn_cuda_vs_reference_golden(cuda):
    from graph_hscn_b200 import models
    from graph_hscn_b200.data import Batch
    g = load("mpnn")
    batch = Batch.from_data_list(graphs_from(g["graphs"])).to(cuda)
    batch.x = batch.x.float()
    m = models.MPNN("gcn", F.relu, 9, 32, 10, 4, ops=_product()).to(cuda)
    m.load_state_dict(g["state"])
    m.eval()
    pred = m(batch)
    loss, score = models.criterion("cross_entropy", pred, batch.y)
    loss.backward()
    assert_close(pred, g["pred"], RTOL, "MPNN pred")
    assert_close(loss, g["loss"], RTOL, "MPNN loss")
    for n, p in m.named_parameters():
        assert_close(p.grad, g["grads"][n], 10 * RTOL, f"MPNN grad {n}")


def test_scn_per_graph_cuda_vs_reference_golden(cuda):
    """The reference's exact call pattern (one graph per call): gcn_norm -> SCN -> to_dense_adj ->
    dense_mincut_pool, weights after the reference's own 2-epoch clustering loop."""
    from graph_hscn_b200 import hetero, models
    g = load("scn")
    p = _product()
    graphs = graphs_from(g["graphs"])
    scn = models.SCN([16], "elu", 9, g["K"], ops=p).to(cuda)
    scn.load_state_dict(g["final_state"])
    d = graphs[0].to(cuda)
    ei, ew = p.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
    assert torch.equal(ei.cpu(), g["g0"]["edge_index"])
    assert torch.equal(ew.cpu(), g["g0"]["edge_weight"])          # gcn_norm weights are bit-exact
    S, mc, ol, adj = scn(d.x.float(), ei, ew)
    assert_close(S, g["g0"]["S"], RTOL, "S")
    assert_close(mc, g["g0"]["mc"], RTOL, "mc")
    assert_close(ol, g["g0"]["ortho"], RTOL, "ortho")
    assert torch.equal(adj.to_dense().cpu(), g["g0"]["adj"])      # lazy adjacency densifies to PyG's tensor
    # cluster ids of every graph: bit-exact wherever the fp32 softmax margin exceeds rounding noise
    agree = total = 0
    for dd, want in zip(graphs, g["clusters"]):
        dd = dd.to(cuda)
        ei, ew = p.gcn_norm(dd.edge_index, None, dd.num_nodes, add_self_loops=True)
        S, _, _, _ = scn(dd.x.float(), ei, ew)
        top2 = S.topk(2, dim=1)[0]
        safe = ((top2[:, 0] - top2[:, 1]) > 1e-5).cpu()
        got = hetero.assign_clusters(S.detach()).cpu().long()
        assert torch.equal(got[safe], want[safe])
        agree += int((got == want).sum())
        total += want.numel()
    assert agree / total > 0.999, f"cluster id agreement {agree}/{total}"


def test_virtual_nodes_cuda_vs_reference_golden(cuda):
    """K7 vs loader/hetero_data.py:42-87 run unchanged: bit-exact features and edge lists."""
    from graph_hscn_b200 import hetero
    from graph_hscn_b200.data import Batch
    g, h = load("scn"), load("hetero")
    graphs = graphs_from(g["graphs"])
    b = Batch.from_data_list(graphs)
    clusters = torch.cat(g["clusters"]).int()
    hb = hetero.build_hetero_batch(b.x.to(cuda), b.edge_index.to(cuda), b.batch.to(cuda), clusters.to(cuda), g["K"])
    voff = eoff = noff = 0
    vx, vv, lv = hb["virtual"].x.cpu(), hb["virtual", "to", "virtual"].edge_index.cpu(), \
        hb["local", "to", "virtual"].edge_index.cpu()
    for want in h:
        U, E, n = want["virtual_x"].size(0), want["vv"].size(1), want["lv"].size(1)
        assert torch.equal(vx[voff:voff + U], want["virtual_x"])
        assert torch.equal(vv[:, eoff:eoff + E] - voff, want["vv"])
        sl = lv[:, noff:noff + n]
        assert torch.equal(sl[0] - noff, want["lv"][0]) and torch.equal(sl[1] - voff, want["lv"][1])
        voff, eoff, noff = voff + U, eoff + E, noff + n
    assert voff == vx.size(0) and eoff == vv.size(1)


def test_hscn_cuda_vs_reference_golden(cuda):
    from graph_hscn_b200 import models
    from graph_hscn_b200.data import Batch, HeteroData
    from torch.nn.parameter import UninitializedParameter
    hg, g = load("hetero"), load("hscn")
    hl = []
    for w in hg:
        h = HeteroData()
        h["local"].x, h["local"].y, h["virtual"].x = w["local_x"], w["y"], w["virtual_x"]
        h["local", "to", "local"].edge_index = w["ll"]
        h["virtual", "to", "virtual"].edge_index = w["vv"]
        h["local", "to", "virtual"].edge_index = w["lv"]
        hl.append(h)
    hb = Batch.from_data_list(hl).to(cuda)
    m = models.HSCN("GAT", "GCN", "GCN", F.relu, 9, 24, 11, 2, ops=_product()).to(cuda)
    for n, prm in m.named_parameters():
        if isinstance(prm, UninitializedParameter):
            prm.materialize(g["state"][n].shape, device=cuda)
    m.load_state_dict(g["state"])
    pred = m(hb.x_dict, hb.edge_index_dict, hb)
    loss, _ = models.criterion("l1", pred, hb["local"].y)
    loss.backward()
    assert_close(pred, g["pred"], RTOL, "HSCN pred")
    assert_close(loss, g["loss"], RTOL, "HSCN loss")
    for n, prm in m.named_parameters():
        want = g["grads"][n]
        if want is None:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, n
        else:
            assert_close(prm.grad, want, 10 * RTOL, f"HSCN grad {n}")
