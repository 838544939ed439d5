"""CPU tests for SURVEY 8f-4 (Laplacian PE + SignNet): the oracle restatement and the mirror encoder against the golden
vectors produced by the reference's own source text (tests/golden/make_golden_posenc.py), plus the properties the
GPU tests rely on."""
import os
import types

import numpy as np
import pytest
import torch

from tests.util import assert_close

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "posenc.pt"), weights_only=False)


def same_with_nan(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and bool(torch.equal(torch.isnan(a), torch.isnan(b))) and bool(
        torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)))


def test_decomp_stats_equal_reference_source_bitwise(gold):
    """oracle.get_lap_decomp_stats / eigvec_normalizer == transform/posenc.py:50-108 on the same LAPACK outputs."""
    from oracle import posenc as op
    assert len(gold["decomp"]) >= 9
    for d in gold["decomp"]:
        vals, vecs = op.get_lap_decomp_stats(d["evals"].numpy(), d["evects"].numpy(), 10, d["vec_norm"])
        assert same_with_nan(vals, d["eigvals_sn"]), (d["n"], d["lap_norm"])
        assert same_with_nan(vecs, d["eigvecs_sn"]), (d["n"], d["lap_norm"])
        assert vals.shape == (d["n"], 10, 1) and vecs.shape == (d["n"], 10)


def test_oracle_laplacian_identities():
    """get_laplacian restatement: symmetric, rows of D - A sum to zero, sym-normalised spectrum inside [0, 2], the
    constant vector (scaled by sqrt(deg)) is in the null space; to_undirected symmetrises and coalesces."""
    from graph_hscn_b200 import synthetic
    from oracle import posenc as op
    g = synthetic.peptides_graphs(1, seed=3, task="func")[0]
    n = g.x.size(0)
    lap = op.laplacian_dense(g.edge_index, n, True, "none").astype(np.float64)
    assert np.allclose(lap, lap.T) and np.allclose(lap.sum(1), 0)
    sym = op.laplacian_dense(g.edge_index, n, True, "sym").astype(np.float64)
    w = np.linalg.eigvalsh(sym)
    assert w.min() > -1e-6 and w.max() < 2 + 1e-6
    deg = np.diag(lap)
    assert np.abs(sym @ np.sqrt(deg)).max() < 1e-5
    half = g.edge_index[:, g.edge_index[0] < g.edge_index[1]]           # one direction only, with a duplicate
    half = torch.cat([half, half[:, :1]], 1)
    assert np.array_equal(op.laplacian_dense(half, n, False, "sym"), op.laplacian_dense(g.edge_index, n, True, "sym"))


def test_oracle_compute_posenc_matches_golden_rows(gold):
    """Full restated pipeline (Laplacian -> eigh -> stats) vs the rows stored with the golden graphs; eigenvector
    columns up to sign, and only where the eigenvalue is isolated (LAPACK's basis inside a cluster is arbitrary)."""
    from oracle import posenc as op
    for g in gold["graphs"]:
        n = g["x"].size(0)
        vals, vecs = op.compute_posenc_stats(g["edge_index"], n, True, 10, "L2", "sym")
        assert torch.equal(torch.isnan(vals), torch.isnan(g["eigvals_sn"]))
        assert torch.allclose(torch.nan_to_num(vals), torch.nan_to_num(g["eigvals_sn"]), atol=2e-6)
        lam = torch.nan_to_num(vals[0, :, 0], nan=1e9)
        for k in range(min(n, 10)):
            gap = min([abs(float(lam[k] - lam[j])) for j in range(10) if j != k])
            if gap > 1e-3 and k < 9:
                dot = float((vecs[:, k] * g["eigvecs_sn"][:, k]).sum())
                assert abs(abs(dot) - 1) < 1e-4, (n, k, dot)


@pytest.mark.parametrize("model", ["DeepSet", "MLP"])
def test_mirror_signnet_equals_reference_source(gold, model):
    """graph_hscn_b200.signnet on the oracle operators == encoder/signnet.py (unmodified source) with the same weights."""
    from graph_hscn_b200 import signnet
    from graph_hscn_b200.data import Batch, Data
    from oracle.namespace import namespace
    e = gold["encoder"][model]
    cfg = types.SimpleNamespace(**e["cfg"])
    enc = signnet.SignNetNodeEncoder(cfg, 9, 24, ops=namespace())
    enc.load_state_dict(e["state"])
    enc.eval()
    graphs = [Data(x=g["x"], edge_index=g["edge_index"], eigvals_sn=g["eigvals_sn"].clone(),
                   eigvecs_sn=g["eigvecs_sn"].clone()) for g in gold["graphs"]]
    batch = Batch.from_data_list(graphs)
    with torch.no_grad():
        out = enc(batch)
    assert_close(out.x, e["x"], 1e-6, f"{model} batch.x")
    assert_close(out.pe_SignNet, e["pe"], 1e-6, f"{model} pe")
    # sign invariance (the property that makes LAPACK's arbitrary signs harmless)
    flip = torch.where(torch.rand(10) < 0.5, -1.0, 1.0)
    graphs2 = [Data(x=g["x"], edge_index=g["edge_index"], eigvals_sn=g["eigvals_sn"].clone(),
                    eigvecs_sn=g["eigvecs_sn"].clone() * flip) for g in gold["graphs"]]
    with torch.no_grad():
        out2 = enc(Batch.from_data_list(graphs2))
    assert_close(out2.pe_SignNet, e["pe"], 1e-5, f"{model} pe under sign flips")


def test_batched_n_nodes_and_gin_depth():
    from graph_hscn_b200 import signnet
    from oracle.namespace import namespace
    b = torch.tensor([0, 0, 0, 1, 2, 2])
    assert signnet.MaskedGINDeepSigns.batched_n_nodes(b).tolist() == [3, 3, 3, 1, 2, 2]
    for n_layers, want in ((1, 2), (2, 2), (3, 3), (4, 4)):              # signnet.py:108-139: first + last always exist
        gin = signnet.GIN(1, 8, 4, n_layers, use_bn=True, ops=namespace())
        assert len(gin.layers) == want and len(gin.bns) == want - 1


def test_signnet_constructor_errors():
    from graph_hscn_b200 import signnet
    from oracle.namespace import namespace
    base = dict(dim_pe=8, layers=2, post_layers=2, eigen_max_freqs=10, phi_hidden_dim=16, phi_out_dim=4,
                pass_as_var=False, use_bn=False)
    with pytest.raises(ValueError):
        signnet.SignNetNodeEncoder(types.SimpleNamespace(model="GNN", **base), 9, 24, ops=namespace())
    with pytest.raises(ValueError):
        signnet.SignNetNodeEncoder(types.SimpleNamespace(model="MLP", **{**base, "post_layers": 0}), 9, 24,
                                   ops=namespace())
    with pytest.raises(ValueError):
        signnet.SignNetNodeEncoder(types.SimpleNamespace(model="MLP", **base), 9, 8, ops=namespace())
    enc = signnet.SignNetNodeEncoder(types.SimpleNamespace(model="MLP", **base), 9, 24, ops=namespace())
    with pytest.raises(ValueError):
        enc(types.SimpleNamespace(x=torch.zeros(2, 9)))


def test_laplacian_eig_argument_errors(built_lib):
    """C-ABI argument validation (no GPU work is launched for a rejected call)."""
    from graph_hscn_b200._lib import lib
    L = lib()
    assert L.query("ghscn_laplacian_eig_workspace_bytes", 100, 4, 50) == (2 * 100 * 50 + 2 * 4) * 8 + 256
    assert L.query("ghscn_laplacian_eig_workspace_bytes", -1, 4, 50) == 0
    fn = L._fns["ghscn_laplacian_eig"]
    one = 8                                                        # fake non-null pointers; rejected before any use
    assert fn(one, one, one, 2, 10, 5, 3, 0, 10, 1, one, one, None, one, 1 << 20, None) == -1   # bad norm
    assert fn(one, one, one, 2, 10, 5, 1, 0, 10, 5, one, one, None, one, 1 << 20, None) == -1   # bad eigvec norm
    assert fn(one, one, one, 2, 10, 5, 1, 0, 0, 1, one, one, None, one, 1 << 20, None) == -1    # max_freqs
    assert fn(one, one, one, 2, 10, 5, 1, 0, 10, 1, one, one, None, one, 16, None) == -2        # workspace
    assert fn(one, one, one, 2, 10, 5000, 1, 0, 10, 1, one, one, None, one, 1 << 40, None) == -3  # n_cap
    assert fn(None, None, None, 0, 0, 0, 1, 0, 10, 1, None, None, None, None, 0, None) == 0     # empty batch


def test_shim_utils_equal_oracle_restatement(gold):
    """graph_hscn_b200/pyg/utils.py (what `torch_geometric.utils` resolves to under pyg.install(), so that
    transform/posenc.py imports and runs unchanged) == the oracle's restatement of the same PyG functions, bit for bit;
    and the golden rows of the reference's is_undirected = False branch."""
    from graph_hscn_b200 import pyg
    from graph_hscn_b200.pyg import utils as pu
    from oracle import posenc as op
    mods = pyg.build_modules(pyg.namespace())
    for name in ("get_laplacian", "to_undirected", "to_scipy_sparse_matrix", "remove_self_loops", "to_dense_adj"):
        assert callable(getattr(mods["torch_geometric.utils"], name)), name
    g = gold["graphs"][0]
    n = g["x"].size(0)
    ei = torch.cat([g["edge_index"], torch.tensor([[3, 3, 0], [3, 3, 1]])], 1)      # self loops + a duplicate
    for norm in (None, "sym", "rw"):
        a_i, a_w = pu.get_laplacian(ei, normalization=norm, num_nodes=n)
        b_i, b_w = op.get_laplacian(ei, normalization=norm, num_nodes=n)
        assert torch.equal(a_i, b_i) and torch.equal(a_w, b_w), norm
        dense_a = pu.to_scipy_sparse_matrix(a_i, a_w, n).toarray()
        dense_b = op.to_scipy_sparse_matrix(b_i, b_w, n).toarray()
        assert dense_a.dtype == np.float32 and np.array_equal(dense_a, dense_b)
    half = ei[:, ei[0] <= ei[1]]
    assert torch.equal(pu.to_undirected(half, num_nodes=n), op.to_undirected(half, n))
    with pytest.raises(ValueError):
        pu.get_laplacian(ei, normalization="bad", num_nodes=n)
    d = gold["directed"]
    vals, vecs = op.compute_posenc_stats(d["edge_index"], d["x"].size(0), False, 10, "L2", "sym")
    assert torch.allclose(vals, d["eigvals_sn"], atol=2e-6)


def test_product_pe_path_is_cuda_only_and_fails_loudly():
    """No CPU fallback: the device eigensolver refuses CPU tensors instead of quietly running numpy."""
    from graph_hscn_b200 import posenc
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError):
        posenc.laplacian_eig(ei, torch.tensor([0, 2]), 2, 2)
    import inspect
    src = inspect.getsource(posenc)
    assert "oracle" not in src.replace("oracle/posenc.py", "") and "numpy" not in src


def test_gin_three_dimensional_input_equals_per_slice(gold):
    """GINConv on the [K, N, C] stack of encoder/signnet.py:227-229 (PyG propagates along node_dim = -2) == the 2-D
    layer applied slice by slice (oracle operators)."""
    from oracle import nn as onn
    g = gold["graphs"][1]
    n = g["x"].size(0)
    torch.manual_seed(4)
    lin = torch.nn.Linear(3, 5)
    conv = onn.GINConv(lin)
    x = torch.randn(4, n, 3)
    full = conv(x, g["edge_index"])
    per = torch.stack([conv(x[k], g["edge_index"]) for k in range(4)])
    assert full.shape == (4, n, 5)
    assert torch.allclose(full, per, atol=1e-6)


def test_oracle_matches_reference_function_for_other_normalisations(gold):
    """compute_posenc_stats of the reference (unmodified) with eigen_laplacian_norm = none / rw and eigvec_norm = L1 /
    abs-max: eigenvalues and, for isolated eigenvalues, eigenvectors up to sign."""
    from oracle import posenc as op
    g = gold["graphs"][1]
    n = g["x"].size(0)
    for row in gold["other_norms"]:
        vals, vecs = op.compute_posenc_stats(g["edge_index"], n, True, 10, row["vec_norm"], row["lap_norm"])
        assert torch.allclose(vals, row["eigvals_sn"], atol=5e-6 * max(1.0, float(row["eigvals_sn"].max())))
        lam = vals[0, :, 0]
        for k in range(9):
            if min(abs(float(lam[k] - lam[j])) for j in range(10) if j != k) > 1e-3:
                a, b = vecs[:, k], row["eigvecs_sn"][:, k]
                sign = 1.0 if float((a * b).sum()) >= 0 else -1.0
                assert float((a - sign * b).abs().max()) < 1e-4, (row["lap_norm"], k)

