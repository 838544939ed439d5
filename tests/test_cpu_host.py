"""CPU tests (no GPU): C-ABI surface, oracle self-checks, collate convention, synthetic shapes, DP host logic."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from tests.util import assert_close, random_edge_index


# ----------------------------------------------------------------------------------------- C ABI
def test_header_and_binding_agree(built_lib):
    from graph_hscn_b200 import _abi_check
    _abi_check.check()
    assert len(_abi_check.header_functions()) >= 25


def test_library_exports_every_declared_symbol(built_lib):
    from graph_hscn_b200 import _abi_check
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = sorted(set(_abi_check.header_functions()) - exported)
    assert not missing, f"declared in include/ghscn.h but not exported: {missing}"
    leaked = sorted(s for s in exported if " T " in s)
    assert not leaked


def test_library_loads_and_validates_arguments_without_a_gpu(built_lib):
    from graph_hscn_b200._lib import GhscnError, lib
    L = lib()
    assert L.query("ghscn_abi_version") == 2
    assert L.error_string(0) == "ok" and "invalid" in L.error_string(-1)
    assert L.query("ghscn_csr_workspace_bytes", 1000, 100, 1) > 3 * 1100 * 4
    # argument errors are reported before anything is launched (negative sizes / null pointers)
    with pytest.raises(GhscnError):
        L.call("ghscn_spmm", None, None, None, None, 4, None, 4, None, -1, 4, 0, None)
    with pytest.raises(GhscnError):
        L.call("ghscn_csr_build", None, None, 10, 4, 0, None, None, None, None, 0, None)
    with pytest.raises(GhscnError):
        L.call("ghscn_mincut_fwd", None, 0, None, 0, None, None, None, None, 1.0, 1, 1, 200, 1, 1,
               None, None, None, None, None, None, None, None, 0, None)


def test_built_for_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_path_refuses_cpu_tensors(built_lib):
    from graph_hscn_b200 import pyg
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pyg.scatter_mean(torch.randn(4, 3), torch.tensor([0, 0, 1, 1]), dim=0)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pyg.gcn_norm(torch.tensor([[0, 1], [1, 0]]))


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(__file__))
    pkg = os.path.join(root, "graph_hscn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


# ------------------------------------------------------------------------------------ oracle checks
def _ops():
    from oracle import ops
    return ops


def test_oracle_gcn_norm_self_loop_rules():
    ops = _ops()
    ei = torch.tensor([[0, 1, 1, 2, 2], [1, 0, 1, 2, 0]])
    w = torch.tensor([1.0, 2.0, 5.0, 7.0, 3.0])
    ei2, w2 = ops.add_remaining_self_loops(ei, w, 1.0, 3)
    assert ei2.tolist() == [[0, 1, 2, 0, 1, 2], [1, 0, 0, 0, 1, 2]]
    assert w2.tolist() == [1.0, 2.0, 3.0, 1.0, 5.0, 7.0]       # existing loops keep their weight
    _, wn = ops.gcn_norm(ei, None, 3, add_self_loops=True)
    deg = torch.tensor([3.0, 2.0, 1.0])                         # in-degree incl. the appended loop
    want = deg.pow(-0.5)[ei2[0]] * deg.pow(-0.5)[ei2[1]]
    assert torch.allclose(wn, want)


def test_oracle_mincut_dense_equals_sparse_identities():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    n, K = 40, 6
    ei = random_edge_index(n, n, 150, g)
    s = torch.randn(n, K, generator=g, dtype=torch.float64)
    x = torch.randn(n, 8, generator=g, dtype=torch.float64)
    adj = ops.to_dense_adj(ei, max_num_nodes=n).double()
    out, out_adj, mc, ol = ops.dense_mincut_pool(x, adj, s)
    S = torch.softmax(s, -1)
    num, den, AS = ops.mincut_sparse_terms(ei, S)
    assert torch.allclose(mc, -num / den)
    assert torch.allclose(AS, adj[0] @ S)
    assert -1.0 - 1e-9 <= float(mc) <= 1e-9 and -1e-9 <= float(ol) <= 2.0 + 1e-9
    assert torch.allclose(out[0], S.t() @ x)
    assert torch.allclose(torch.diagonal(out_adj[0]), torch.zeros(K, dtype=torch.float64))


def test_oracle_mincut_analytic_backward_matches_autograd():
    """SURVEY A.7 (the formulas the CUDA backward implements), checked in fp64 on the oracle."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    n, K = 30, 5
    ei = random_edge_index(n, n, 100, g)
    z = torch.randn(n, K, generator=g, dtype=torch.float64, requires_grad=True)
    adj = ops.to_dense_adj(ei, max_num_nodes=n).double()
    _, _, mc, ol = ops.dense_mincut_pool(torch.zeros(n, 1, dtype=torch.float64), adj, z)
    (mc + ol).backward()
    A = adj[0]
    S = torch.softmax(z.detach(), -1)
    d = A.sum(-1)
    num, den = (S * (A @ S)).sum(), (d[:, None] * S * S).sum()
    dS = -(((A @ S) + (A.t() @ S)) * den - num * 2 * d[:, None] * S) / den ** 2
    SS = S.t() @ S
    Fn = SS.norm()
    M = SS / Fn
    R = M - torch.eye(K, dtype=torch.float64) / math.sqrt(K)
    G = R / R.norm()
    Gp = (G - M * (G * M).sum()) / Fn
    dS = dS + S @ (Gp + Gp.t())
    dz = S * (dS - (dS * S).sum(-1, keepdim=True))
    assert torch.allclose(dz, z.grad, rtol=1e-9, atol=1e-12)


def test_oracle_gcnconv_permutation_equivariance_and_gat_softmax():
    from oracle import nn as onn
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    n = 50
    ei = random_edge_index(n, n, 200, g)
    x = torch.randn(n, 7, generator=g)
    torch.manual_seed(0)
    conv = onn.GCNConv(7, 5)
    perm = torch.randperm(n, generator=g)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    y = conv(x, ei)
    y_p = conv(x[perm], inv[ei])
    assert torch.allclose(y[perm], y_p, atol=1e-5)
    a = ops.segment_softmax(torch.randn(200, generator=g), ei[1], n)
    sums = ops.scatter_sum(a, ei[1], 0, None, n)
    has = torch.bincount(ei[1], minlength=n) > 0
    assert torch.allclose(sums[has], torch.ones(int(has.sum())), atol=1e-5)


def test_oracle_scatter_mean_and_pool():
    ops = _ops()
    x = torch.tensor([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    idx = torch.tensor([0, 0, 2])
    out = ops.scatter_mean(x, idx, dim=0)
    assert out.tolist() == [[2.0, 3.0], [0.0, 0.0], [5.0, 6.0]]
    assert torch.equal(ops.global_mean_pool(x, idx), out)


# --------------------------------------------------------------------------------- data / collate
def test_collate_batch_ptr_convention():
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    graphs = synthetic.peptides_graphs(5, seed=9)
    b = Batch.from_data_list(graphs)
    n = [g.num_nodes for g in graphs]
    assert b.ptr.tolist() == np.concatenate([[0], np.cumsum(n)]).tolist()
    assert bool((b.batch[1:] >= b.batch[:-1]).all()) and b.num_graphs == 5
    off = 0
    for gi, g in enumerate(graphs):
        e = g.edge_index.size(1)
        assert torch.equal(b.edge_index[:, off:off + e], g.edge_index + int(b.ptr[gi]))
        off += e
    assert b.y.shape == (5, 10)
    back = b.to_data_list()
    assert all(torch.equal(a.edge_index, c.edge_index) and torch.equal(a.x, c.x) for a, c in zip(back, graphs))
    assert graphs[0].edge_weight is None          # PyG: missing standard attributes read as None


def test_hetero_collate_offsets():
    from graph_hscn_b200.data import Batch, HeteroData
    hl = []
    for n, u in [(4, 2), (3, 1)]:
        h = HeteroData()
        h["local"].x = torch.zeros(n, 2)
        h["local"].y = torch.zeros(1, 3)
        h["virtual"].x = torch.zeros(u, 2)
        h["local", "to", "local"].edge_index = torch.tensor([[0, 1], [1, 0]])
        h["virtual", "to", "virtual"].edge_index = torch.tensor([[0], [0]])
        h["local", "to", "virtual"].edge_index = torch.stack([torch.arange(n), torch.zeros(n, dtype=torch.long)])
        hl.append(h)
    b = Batch.from_data_list(hl)
    assert list(b.edge_index_dict) == [("local", "to", "local"), ("virtual", "to", "virtual"),
                                       ("local", "to", "virtual")]
    assert b["local", "to", "local"].edge_index.tolist() == [[0, 1, 4, 5], [1, 0, 5, 4]]
    assert b["virtual", "to", "virtual"].edge_index.tolist() == [[0, 2], [0, 2]]
    lv = b["local", "to", "virtual"].edge_index
    assert lv[0].tolist() == list(range(7)) and lv[1].tolist() == [0, 0, 0, 0, 2, 2, 2]
    assert b["local"].batch.tolist() == [0, 0, 0, 0, 1, 1, 1] and b["virtual"].batch.tolist() == [0, 0, 1]
    assert b["local"].y.shape == (2, 3)


def test_dataloader_and_synthetic_shapes():
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import DataLoader
    graphs = synthetic.peptides_graphs(256, seed=1234)
    n = np.array([g.num_nodes for g in graphs])
    e = np.array([g.num_edges for g in graphs])
    assert 135 < n.mean() < 168 and n.min() >= 8 and n.max() <= 444
    assert 1.95 < (e / n).mean() < 2.12                     # ~307 directed edges per ~151 nodes
    g0 = graphs[0]
    assert g0.x.dtype == torch.int64 and g0.x.shape[1] == 9 and int(g0.x[:, 0].max()) < 119
    assert torch.equal(g0.edge_index[:, 0::2], g0.edge_index[:, 1::2].flip(0))   # (i,j),(j,i) adjacent
    deg = torch.bincount(g0.edge_index[0], minlength=g0.num_nodes)
    assert int(deg.max()) <= 6
    loader = DataLoader(graphs, 32, shuffle=False, num_workers=0, persistent_workers=False)
    b = next(iter(loader))
    assert b.num_graphs == 32 and len(loader) == 8
    v = synthetic.vocsp_graphs(8)
    dv = np.mean([g.num_edges / g.num_nodes for g in v])
    assert 4.5 < dv < 6.5 and v[0].x.shape[1] == 14 and 395 <= v[0].num_nodes <= 500


def test_shard_range_covers_all_graphs():
    from graph_hscn_b200.train import shard_range
    for B, W in [(128, 8), (10, 4), (3, 8), (1024, 3)]:
        r = [shard_range(B, k, W) for k in range(W)]
        assert r[0][0] == 0 and r[-1][1] == B and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
    w = [1] * 50 + [9] * 50
    r = [shard_range(100, k, 2, w) for k in range(2)]
    assert r[0][1] > 50 and r[0][1] == r[1][0] and r[1][1] == 100


def test_gemm3x_shape_rules_without_a_gpu(built_lib):
    """ghscn_gemm3x_supported / *_bytes are host-only: shape rules of the tcgen05 GEMMs (include/ghscn.h)."""
    from graph_hscn_b200._lib import lib
    L = lib()
    assert L.query("ghscn_gemm3x_supported", 18269, 300, 300) == 1
    assert L.query("ghscn_gemm3x_supported", 18269, 320, 300) == 1
    assert L.query("ghscn_gemm3x_supported", 18269, 324, 300) == 0      # more than two 160-column halves
    assert L.query("ghscn_gemm3x_supported", 18269, 302, 300) == 0      # n_out % 4
    assert L.query("ghscn_gemm3x_supported", 18269, 300, 9) == 0        # k % 4 (raw atom features: skinny kernels)
    assert L.query("ghscn_gemm3x_supported", 0, 300, 300) == 0
    # weight image: per N half, per 32-wide K chunk, hi + lo blocks of pad x 128 bytes; 300 -> halves 160 + 144
    assert L.query("ghscn_gemm3x_b_image_bytes", 300, 300) == 10 * 2 * (160 + 144) * 128
    assert L.query("ghscn_gemm3x_b_image_bytes", 64, 64) == 2 * 2 * 64 * 128
    assert L.query("ghscn_gemm3x_b_image_bytes", 400, 300) == 0
    assert L.query("ghscn_gemm3x_tn_supported", 18269, 300, 300) == 1
    assert L.query("ghscn_gemm3x_tn_supported", 18269, 300, 324) == 0
    ws = L.query("ghscn_gemm3x_tn_workspace_bytes", 18269, 300, 300)
    assert ws % (300 * 300 * 4) == 0 and 1 <= ws // (300 * 300 * 4) <= 64   # one fp32 partial per row slab
    # null pointers are rejected before anything is launched
    assert L._fns["ghscn_gemm3x"](None, 300, 128, 300, None, 300, None, 0, None, 300, None) == -1
    assert L._fns["ghscn_gemm3x_tn"](None, 300, None, 300, 128, 300, 300, None, None, 0, None) == -1


def test_edge_blocks_from_batch_accepts_only_collated_edge_lists():
    """Host-side gate of the K1 fast path: graph-major, block-diagonal edge lists only."""
    import torch
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.structure import edge_blocks_from_batch
    b = synthetic.peptides_batch(6, seed=3)
    counts = b.ptr[1:] - b.ptr[:-1]
    got = edge_blocks_from_batch(b.edge_index, b.batch, 6)
    assert got == (int(counts.max()), int(torch.bincount(b.batch[b.edge_index[0]], minlength=6).max()))
    crossing = b.edge_index.clone()
    crossing[1, 0] = int(b.ptr[3])                      # first edge now ends in graph 3
    assert edge_blocks_from_batch(crossing, b.batch, 6) is None
    assert edge_blocks_from_batch(b.edge_index.flip(1), b.batch, 6) is None      # graph-minor order
    assert edge_blocks_from_batch(b.edge_index[:, :0], b.batch, 6) is None


def test_collate_records_block_bounds():
    """Batch.from_data_list keeps the collate-time facts the per-graph CSR kernel needs (no device sync later)."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    graphs = synthetic.peptides_graphs(7, seed=2)
    b = Batch.from_data_list(graphs)
    assert b.max_nodes_per_graph == max(g.num_nodes for g in graphs)
    assert b.max_edges_per_graph == max(g.edge_index.size(1) for g in graphs)
    back = b.to_data_list()
    assert len(back) == 7 and all("max_edges_per_graph" not in g for g in back)
    assert Batch.from_data_list(back).max_edges_per_graph == b.max_edges_per_graph


def test_collate_refuses_block_bounds_for_edges_leaving_their_graph():
    """A Data object whose edge index points outside its own node range must not enable the per-graph CSR kernel."""
    import torch
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    graphs = synthetic.peptides_graphs(4, seed=5)
    graphs[1].edge_index = graphs[1].edge_index.clone()
    graphs[1].edge_index[1, 0] = graphs[1].num_nodes + 3          # lands in the next graph after collate
    b = Batch.from_data_list(graphs)
    assert "max_edges_per_graph" not in b and b.max_nodes_per_graph > 0


def test_flat_gradients_backward_into_equals_backward():
    """FlatGradients.backward_into (autograd.grad + one multi-tensor copy) == zero the flat buffer + loss.backward()."""
    import torch
    from graph_hscn_b200.train import FlatGradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    shared = torch.nn.Linear(3, 3)                       # used twice: its gradient has two contributions
    mod = torch.nn.ModuleList([net, shared])
    fg = FlatGradients(mod)
    x = torch.randn(11, 5)

    def loss():
        y = net(x)
        return (shared(shared(y)) ** 2).mean()
    fg.flat.fill_(123.0)                                 # stale content must be overwritten, not accumulated onto
    fg.backward_into(loss())
    got = fg.flat.clone()
    fg.zero()
    loss().backward()
    assert torch.equal(got, fg.flat)
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in fg.params)     # still views of the flat buffer


def test_new_entry_points_validate_arguments_without_a_gpu(built_lib):
    """Per-graph CSR build, fused SCN pipeline, segmented tcgen05 contraction, split MinCUT forward: argument and
    shape errors are reported before any launch (runs on the CPU-only build box)."""
    from graph_hscn_b200._lib import GhscnError, lib
    L = lib()
    assert L.query("ghscn_csr_blocked_smem_bytes", 444, 920) == (4 * 444 + 2 + 4 * 920) * 4
    assert L.query("ghscn_scn_backward_workspace_bytes", 18269, 9, 16, 10) == 143 * 474 * 4
    with pytest.raises(GhscnError, match="invalid"):
        L.call("ghscn_csr_build_blocked", None, None, 10, None, 1, 10, 10, 10, None, None, None, None, None, None,
               None, None)
    with pytest.raises(GhscnError, match="unsupported"):      # 17 input features: outside the fused kernel's range
        L.call("ghscn_scn_forward", None, None, None, None, 17, 100, 17, 16, 10, None, None, None, None, None, 1,
               None, None, None, None, None)
    with pytest.raises(GhscnError, match="unsupported"):      # activation id out of range
        L.call("ghscn_scn_backward", None, None, None, None, None, 9, 100, 9, 16, 10, None, 7, None, None, 0, None)
    with pytest.raises(GhscnError, match="invalid"):
        L.call("ghscn_gemm3x_tn_segmented", None, 64, None, 300, None, 4, 100, 64, 300, None, 300, 64 * 300, None)
    with pytest.raises(GhscnError, match="invalid"):          # phase out of range
        L.call("ghscn_mincut_fwd_phase", None, 0, None, 0, None, None, None, None, 1.0, 1, 1, 10, 1, 1,
               None, None, None, None, None, None, None, None, 0, None, 4)
    # round 2: split MinCUT backward, fused output layer + loss, fused ReLU-backward + column sums
    assert L.query("ghscn_mincut_bwd_split_supported", 128, 512, 512, 512, 444) == 1
    assert L.query("ghscn_mincut_bwd_split_supported", 10, 16, 16, 16, 444) == 0       # K % 4 != 0
    assert L.query("ghscn_mincut_bwd_split_supported", 64, 301, 301, 301, 444) == 0    # rows not 16-byte aligned
    assert L.query("ghscn_mincut_bwd_split_workspace_bytes", 1000, 8, 64) == (4 * 1000 * 64 + 3 * 8 * 64 * 64) * 4 + 256
    with pytest.raises(GhscnError, match="unsupported"):
        L.call("ghscn_mincut_bwd_split", None, None, 0, None, None, None, None, None, None, None, 1.0, 4, 100, 10, 16,
               50, None, None, None, None, None, None, None, 10, None, 16, None, 0, None)
    with pytest.raises(GhscnError, match="workspace"):
        L.call("ghscn_mincut_bwd_split", ctypes.c_void_p(16), None, 0, ctypes.c_void_p(16), ctypes.c_void_p(16), None,
               None, ctypes.c_void_p(16), None, None, 1.0, 4, 100, 64, 16, 50, ctypes.c_void_p(16), ctypes.c_void_p(16),
               ctypes.c_void_p(16), None, None, None, ctypes.c_void_p(16), 64, None, 16, None, 0, None)
    assert L.query("ghscn_head_out_loss_supported", 130, 300, 10) == 1
    assert L.query("ghscn_head_out_loss_supported", 1027, 300, 11) == 0
    with pytest.raises(GhscnError, match="unsupported"):
        L.call("ghscn_head_out_loss", None, 300, None, 300, None, None, 10, 1000, 1000, 300, 10, 0, None, None, None,
               None, None, None, None)
    with pytest.raises(GhscnError, match="invalid"):          # loss mode out of range
        L.call("ghscn_head_out_loss", None, 300, None, 300, None, None, 10, 100, 100, 300, 10, 2, None, None, None,
               None, None, None, None)
    with pytest.raises(GhscnError, match="invalid"):          # masked output without a mask
        L.call("ghscn_relu_grad_colsum_partial", ctypes.c_void_p(16), 8, None, 0, 4, 8, ctypes.c_void_p(16), 8,
               ctypes.c_void_p(16), 1 << 20, None)
    with pytest.raises(GhscnError, match="workspace"):
        L.call("ghscn_colsum_finish", None, 0, 100, 8, ctypes.c_void_p(16), None)
    assert L.query("ghscn_grad_clip_workspace_bytes", 277_000) == 148 * 4
    with pytest.raises(GhscnError, match="workspace"):        # clip_grad_norm: workspace too small
        L.call("ghscn_grad_clip_scale", None, 0, 1.0, None, 0, ctypes.c_void_p(8), None)
    with pytest.raises(GhscnError, match="invalid"):          # max_norm must be positive
        L.call("ghscn_grad_clip_scale", None, 0, 0.0, ctypes.c_void_p(8), 1024, ctypes.c_void_p(8), None)
    with pytest.raises(GhscnError, match="invalid"):          # scaled AdamW without a step counter
        L.call("ghscn_adamw_step_scaled", None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 0.0, None, None, None)


# ------------------------------------------------------------------------------------------------ round 2 host logic
def test_bucket_padding_builds_wellformed_dummy_graphs():
    """train.stage_batch pads a batch into its bucket with trailing dummy graphs: still sorted `batch`, block-diagonal
    graph-major loop-free edges, per-graph caps respected, real rows untouched."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.structure import edge_blocks_from_batch
    from graph_hscn_b200.train import BucketPolicy, stage_batch
    batches = [synthetic.peptides_batch(12, seed=s) for s in (1, 2, 3, 4)]
    pol = BucketPolicy.for_batches(batches, node_step=64, edge_step=128)
    seen = set()
    for b in batches:
        st = stage_batch(b, pol)
        sh, v = st.shape, st.views
        N, E = b.x.size(0), b.edge_index.size(1)
        assert sh.n_cap % 64 == 0 and sh.n_cap >= N + 2 and sh.n_cap - N <= 2 * 64 + 2
        assert sh.e_cap - E <= pol.max_pad_degree * (sh.n_cap - N)
        assert sh.e_cap % pol.eff_edge_step == 0 and 0 <= sh.e_cap - E < pol.eff_edge_step
        assert sh.graphs == 12 and sh.dummies == pol.dummy_graphs
        assert torch.equal(v["x"][:N], b.x) and torch.equal(v["edge_index"][:, :E], b.edge_index)
        assert torch.equal(v["batch"][:N], b.batch) and torch.equal(v["y"][:12], b.y)
        assert not v["x"][N:].any() and not v["y"][12:].any()
        bt = v["batch"]
        assert bool((bt[1:] >= bt[:-1]).all()) and int(bt[-1]) <= 12 + sh.dummies - 1 and int(bt[N]) >= 12
        ei = v["edge_index"]
        assert int(ei.min()) >= 0 and int(ei.max()) < sh.n_cap and not bool((ei[0] == ei[1]).any())
        blocks = edge_blocks_from_batch(ei, bt, 12 + sh.dummies)
        assert blocks is not None and blocks[0] <= sh.max_nodes and blocks[1] <= sh.max_edges
        seen.add((sh.n_cap, sh.e_cap))
    assert len(seen) >= 2                                   # different sizes land in different buckets
    exact = stage_batch(batches[0], None)
    assert exact.shape.dummies == 0 and exact.shape.n_cap == batches[0].x.size(0)


def test_dummy_graphs_do_not_change_losses_or_gradients():
    """The padded batch (B real + D dummy graphs, loss over the first B graphs) gives the gradients of the unpadded
    batch: run on the CPU oracle operators with the MPNN as the local chain of the HSCN."""
    from graph_hscn_b200 import models, synthetic
    from graph_hscn_b200.data import Batch
    from graph_hscn_b200.train import BucketPolicy, stage_batch
    from oracle.namespace import namespace
    b = synthetic.peptides_batch(6, seed=9)
    st = stage_batch(b, BucketPolicy.for_batches([b], node_step=32, edge_step=64))
    v = st.views
    torch.manual_seed(0)
    m = models.MPNN("gcn", torch.relu, 9, 24, 10, 3, ops=namespace())
    b.x = b.x.float()
    loss, _ = models.criterion("cross_entropy", m(b), b.y)
    want = torch.autograd.grad(loss, list(m.parameters()))
    pb = Batch(x=v["x"].float(), edge_index=v["edge_index"], y=v["y"])
    pb.batch = v["batch"]
    pred = m(pb)
    assert pred.size(0) == 6 + st.shape.dummies
    loss_p, _ = models.criterion("cross_entropy", pred[:6], v["y"][:6])
    got = torch.autograd.grad(loss_p, list(m.parameters()))
    assert abs(loss_p.item() - loss.item()) <= 1e-6 * abs(loss.item())
    for a, c in zip(got, want):
        assert float((a - c).abs().max()) <= 1e-6 * float(c.abs().max().clamp_min(1e-12))


def test_balanced_partition_equal_counts_and_sizes():
    from graph_hscn_b200.train import balanced_partition
    rng = np.random.default_rng(0)
    sizes = np.clip(np.rint(rng.normal(151, 60, size=1024)), 8, 444).astype(int).tolist()
    groups = balanced_partition(sizes, 8)
    assert sorted(i for g in groups for i in g) == list(range(1024))
    assert all(len(g) == 128 for g in groups)
    sums = [sum(sizes[i] for i in g) for g in groups]
    assert (max(sums) - min(sums)) / (sum(sums) / 8) < 0.01          # node counts within 1 % across ranks
    uneven = balanced_partition(sizes[:10], 4)
    assert sorted(len(g) for g in uneven) == [2, 2, 3, 3]


def test_device_staging_plan_without_a_gpu():
    """pyg/_device.py: strict mode raises on CPU tensors; auto mode plans a CUDA staging (no CPU fallback)."""
    from graph_hscn_b200.pyg import _device
    x = torch.zeros(3, 2)
    assert not _device.auto_device()
    with pytest.raises(RuntimeError, match="CUDA-only"):
        _device.plan(x)
    _device.set_auto_device(True)
    try:
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                _device.plan(x)
    finally:
        _device.set_auto_device(False)
