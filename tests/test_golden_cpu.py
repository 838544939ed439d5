"""CPU tests: golden vectors produced by the UNMODIFIED reference sources (tests/golden/make_golden.py)
vs this repo's mirror models / restatements running on the oracle operator set."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)


def graphs_from(blob):
    from graph_hscn_b200.data import Data
    return [Data(x=g["x"], edge_index=g["edge_index"], y=g["y"]) for g in blob]


def oracle_ns():
    from oracle.namespace import namespace
    return namespace()


def test_mirror_mpnn_equals_reference_source():
    from graph_hscn_b200 import models
    from graph_hscn_b200.data import Batch
    g = load("mpnn")
    batch = Batch.from_data_list(graphs_from(g["graphs"]))
    batch.x = batch.x.float()
    m = models.MPNN("gcn", F.relu, 9, 32, 10, 4, ops=oracle_ns())
    m.load_state_dict(g["state"])
    m.eval()
    pred = m(batch)
    loss, score = models.criterion("cross_entropy", pred, batch.y)
    loss.backward()
    assert torch.equal(pred, g["pred"]) and torch.equal(loss, g["loss"]) and torch.equal(score, g["score"])
    for n, p in m.named_parameters():
        assert torch.equal(p.grad, g["grads"][n]), n


def test_mirror_scn_and_clustering_loop_equal_reference_source():
    """Replays train_clustering.py:34-69 with the mirror SCN: identical weights and cluster ids."""
    from graph_hscn_b200 import models
    from oracle import hetero as ohet
    g = load("scn")
    o = oracle_ns()
    graphs = graphs_from(g["graphs"])
    scn = models.SCN([16], "elu", 9, g["K"], ops=o)
    scn.load_state_dict(g["init_state"])
    opt = torch.optim.AdamW(scn.parameters(), lr=0.01, weight_decay=5e-4)
    for _ in range(2):
        for d in graphs:
            ei, ew = o.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
            opt.zero_grad()
            _, mc, ol, _ = scn(d.x.float(), ei, ew)
            (mc + ol).backward()
            opt.step()
    for k, v in scn.state_dict().items():
        assert torch.equal(v, g["final_state"][k]), k
    for d, want in zip(graphs, g["clusters"]):
        ei, ew = o.gcn_norm(d.edge_index, None, d.num_nodes, add_self_loops=True)
        S, _, _, _ = scn(d.x.float(), ei, ew)
        assert np.array_equal(ohet.cluster_argmax(S), want.numpy())
    g0 = g["g0"]
    S, mc, ol, adj = scn(graphs[0].x.float(), g0["edge_index"], g0["edge_weight"])
    assert torch.equal(S, g0["S"]) and torch.equal(mc, g0["mc"]) and torch.equal(ol, g0["ortho"])
    assert torch.equal(adj, g0["adj"])


def test_virtual_node_restatement_equals_reference_source():
    """oracle/hetero.py vs loader/hetero_data.py:42-87 run unchanged (split order train, val, test)."""
    from oracle import hetero as ohet
    g, h = load("scn"), load("hetero")
    graphs = graphs_from(g["graphs"])
    assert len(h) == len(graphs)
    for d, c, want in zip(graphs, g["clusters"], h):
        _, vx, vv, lv = ohet.virtual_nodes(d.x, c.numpy(), g["K"])
        assert torch.equal(vx, want["virtual_x"]) and torch.equal(vv, want["vv"]) and torch.equal(lv, want["lv"])
        assert torch.equal(d.x.float(), want["local_x"]) and torch.equal(d.edge_index, want["ll"])


def test_mirror_hscn_equals_reference_source():
    from graph_hscn_b200 import models
    from graph_hscn_b200.data import Batch, HeteroData
    hg, g = load("hetero"), load("hscn")
    hl = []
    for w in hg:
        h = HeteroData()
        h["local"].x, h["local"].y, h["virtual"].x = w["local_x"], w["y"], w["virtual_x"]
        h["local", "to", "local"].edge_index = w["ll"]
        h["virtual", "to", "virtual"].edge_index = w["vv"]
        h["local", "to", "virtual"].edge_index = w["lv"]
        hl.append(h)
    hb = Batch.from_data_list(hl)
    m = models.HSCN("GAT", "GCN", "GCN", F.relu, 9, 24, 11, 2, ops=oracle_ns())
    m(hb.x_dict, hb.edge_index_dict, hb)          # materialise lazy parameters
    m.load_state_dict(g["state"])
    pred = m(hb.x_dict, hb.edge_index_dict, hb)
    loss, _ = models.criterion("l1", pred, hb["local"].y)
    loss.backward()
    assert torch.equal(pred, g["pred"]) and torch.equal(loss, g["loss"])
    for n, p in m.named_parameters():
        want = g["grads"][n]
        if want is None:
            assert p.grad is None, n      # l->v and v->v branches are dead w.r.t. the loss
        else:
            assert torch.equal(p.grad, want), n
    dead = [n for n, w in g["grads"].items() if w is None]
    assert any("local__to__virtual" in n for n in dead) and any("virtual__to__virtual" in n for n in dead)


def test_reference_sources_run_unchanged_on_the_shim():
    """Only where /root/reference exists (the build container): import and run the reference models on the
    oracle shim again and compare with the committed fixtures (guards make_golden.py against drift)."""
    if not os.path.isdir("/root/reference/graph_hscn"):
        pytest.skip("/root/reference not present (GPU box)")
    import subprocess
    import sys
    code = ("import sys; sys.argv=['x']; sys.path.insert(0, %r); import importlib.util as u;"
            "s=u.spec_from_file_location('mg', %r); m=u.module_from_spec(s); s.loader.exec_module(m);"
            "m.install_reference_imports();"
            "import torch, torch.nn.functional as F;"
            "from graph_hscn.model.mpnn import MPNN; from torch_geometric.nn import GCNConv;"
            "from graph_hscn_b200.data import Batch, Data;"
            "g=torch.load(%r, weights_only=False);"
            "b=Batch.from_data_list([Data(x=d['x'],edge_index=d['edge_index'],y=d['y']) for d in g['graphs']]);"
            "b.x=b.x.float(); mm=MPNN(GCNConv,F.relu,9,32,10,4,dropout=0.0); mm.load_state_dict(g['state']);"
            "mm.eval(); assert torch.equal(mm(b), g['pred']); print('ok')"
            ) % (os.path.dirname(os.path.dirname(__file__)), os.path.join(GOLD, "make_golden.py"),
                 os.path.join(GOLD, "mpnn.pt"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
