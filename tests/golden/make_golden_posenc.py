"""Generates tests/golden/posenc.pt from the UNMODIFIED reference sources (build container only; /root/reference does
not exist on the GPU box):

    python tests/golden/make_golden_posenc.py

Pinned here
  * graph_hscn/transform/posenc.py: `get_lap_decomp_stats` and `eigvec_normalizer` -- their source text is cut out of
    the file with `ast` and executed as is on explicit LAPACK outputs; and the whole module imported unchanged, its
    `compute_posenc_stats` run per graph on top of the oracle's restatement of the three PyG utilities it imports
    (`get_laplacian`, `to_scipy_sparse_matrix`, `to_undirected`: torch_geometric is not installable offline, so those
    three stay "parity unpinned") -- the stored rows are asserted to be exactly what that function produces;
  * graph_hscn/encoder/signnet.py: imported unchanged on top of the oracle operator namespace.  `MLP.__init__` reads
    `ACT_DICT["activation"]` (signnet.py:49), a key the reference's dictionary does not have, so
    the generator adds that ONE key (-> relu, the value every call site passes) to the imported dictionary; no source
    line is changed.
The Laplacians come from oracle/posenc.laplacian_dense and `np.linalg.eigh` exactly as posenc.py:30-42 calls them.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def reference_functions(path: str, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "F": F}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def main() -> None:
    from make_golden import install_reference_imports
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch, Data
    from oracle import posenc as op
    get_stats, _norm = reference_functions(os.path.join(REFERENCE, "graph_hscn/transform/posenc.py"),
                                           ["get_lap_decomp_stats", "eigvec_normalizer"])
    out = {"decomp": [], "graphs": []}
    torch.manual_seed(404)                              # the hand-made path graph below draws its features
    graphs = synthetic.peptides_graphs(5, seed=404, task="func")
    path6 = torch.tensor([[0, 1, 1, 2, 2, 3, 3, 4, 4, 5], [1, 0, 2, 1, 3, 2, 4, 3, 5, 4]])
    graphs.append(Data(x=torch.randint(0, 5, (6, 9)), edge_index=path6, y=torch.zeros(1, 10)))   # n < max_freqs
    for gi, g in enumerate(graphs):
        n = g.x.size(0)
        # full LAPACK outputs are kept for three graphs only (fixture size); the others carry the default config's rows
        combos = (("sym", "L2"), ("none", "L1"), ("rw", "abs-max")) if gi in (0, 1, len(graphs) - 1) else (("sym", "L2"),)
        for lap_norm, vec_norm in combos:
            lap = op.laplacian_dense(g.edge_index, n, True, lap_norm)
            evals, evects = np.linalg.eigh(lap)
            vals, vecs = get_stats(evals=evals, evects=evects, max_freqs=10, eigvec_norm=vec_norm)
            if lap_norm == "sym":
                g.eigvals_sn, g.eigvecs_sn = vals.clone(), vecs.clone()
            if len(combos) == 3:
                out["decomp"].append(dict(graph=gi, n=n, lap_norm=lap_norm, vec_norm=vec_norm,
                                          evals=torch.from_numpy(evals.copy()), evects=torch.from_numpy(evects.copy()),
                                          eigvals_sn=vals, eigvecs_sn=vecs))
        out["graphs"].append(dict(x=g.x, edge_index=g.edge_index, eigvals_sn=g.eigvals_sn, eigvecs_sn=g.eigvecs_sn))

    # ---- the encoder, reference source on the oracle operators -----------------------------------------------------
    install_reference_imports()
    import graph_hscn.config.config as ref_config
    ref_config.ACT_DICT["activation"] = F.relu          # see the module docstring
    # ---- transform/posenc.py imported whole and run unmodified on the oracle's get_laplacian / to_undirected /
    #      to_scipy_sparse_matrix: the rows stored above must be what compute_posenc_stats itself produces
    from graph_hscn.transform.posenc import compute_posenc_stats as ref_compute
    pe_cfg = types.SimpleNamespace(eigen_laplacian_norm="sym", eigen_max_freqs=10, eigvec_norm="L2")
    for g in graphs:
        d = ref_compute(Data(x=g.x, edge_index=g.edge_index), True, pe_cfg)
        assert torch.equal(torch.nan_to_num(d.eigvals_sn), torch.nan_to_num(g.eigvals_sn))
        assert torch.equal(torch.nan_to_num(d.eigvecs_sn), torch.nan_to_num(g.eigvecs_sn))
    out["other_norms"] = []
    for lap_norm, vec_norm in (("none", "L1"), ("rw", "abs-max")):        # the non-default PEConfig choices, whole function
        cfg2 = types.SimpleNamespace(eigen_laplacian_norm=lap_norm, eigen_max_freqs=10, eigvec_norm=vec_norm)
        d = ref_compute(Data(x=graphs[1].x, edge_index=graphs[1].edge_index), True, cfg2)
        out["other_norms"].append(dict(lap_norm=lap_norm, vec_norm=vec_norm, eigvals_sn=d.eigvals_sn.clone(),
                                       eigvecs_sn=d.eigvecs_sn.clone()))
    und = graphs[0].edge_index[:, graphs[0].edge_index[0] < graphs[0].edge_index[1]]
    d = ref_compute(Data(x=graphs[0].x, edge_index=und), False, pe_cfg)          # is_undirected = False branch
    out["directed"] = dict(x=graphs[0].x, edge_index=und, eigvals_sn=d.eigvals_sn, eigvecs_sn=d.eigvecs_sn)
    from graph_hscn.encoder.signnet import SignNetNodeEncoder
    base = dict(dim_pe=8, layers=2, post_layers=2, eigen_max_freqs=10, phi_hidden_dim=16, phi_out_dim=4,
                pass_as_var=True, use_bn=False)
    out["encoder"] = {}
    for model in ("DeepSet", "MLP"):
        torch.manual_seed(55)
        cfg = types.SimpleNamespace(model=model, **base)
        enc = SignNetNodeEncoder(cfg, 9, 24)
        enc.eval()
        batch = Batch.from_data_list([g.clone() for g in graphs])
        with torch.no_grad():
            res = enc(batch)
        out["encoder"][model] = dict(cfg=dict(model=model, **base), state={k: v.clone() for k, v in enc.state_dict().items()},
                                     x=res.x.clone(), pe=res.pe_SignNet.clone())
    torch.save(out, os.path.join(HERE, "posenc.pt"))
    print("wrote posenc.pt:", len(out["decomp"]), "decompositions,", {k: tuple(v["x"].shape) for k, v in out["encoder"].items()})


if __name__ == "__main__":
    main()
