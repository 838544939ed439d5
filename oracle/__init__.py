"""CPU oracle for the Graph-HSCN hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain CPU PyTorch / numpy, the arithmetic that the
reference (camille-004/Graph-HSCN) obtains from its un-vendored dependency
PyTorch Geometric (unpinned; torch==1.13.1 era => PyG 2.1-2.3) and
torch_scatter, for exactly the call sites listed in SURVEY.md section 8a:

    graph_hscn/model/mpnn.py:49-60          GCNConv stack + scatter_mean readout
    graph_hscn/model/hscn.py:30-64          GraphConv/Sequential/Linear + to_dense_adj + dense_mincut_pool
    graph_hscn/model/hscn.py:83-125         HeteroConv{GAT l->v, GCN l->l, GCN v->v} + global_mean_pool
    graph_hscn/train/train_clustering.py:34-69   gcn_norm(add_self_loops=True), softmax/argmax cluster ids
    graph_hscn/loader/hetero_data.py:42-87  cluster -> virtual-node construction

PARITY STATUS: *** parity unpinned ***.  The reference ships no tests, golden
vectors or fixtures for this path and PyG / torch_scatter are not installable
in the build container (no network, no wheels).  The restatement follows the
published PyG 2.2/2.3 algorithms and is anchored on the
reference's own call sites: tests/golden/make_golden.py imports the UNMODIFIED
reference sources (graph_hscn/model/*.py, train/train_clustering.py,
loader/hetero_data.py) from /root/reference, runs them on top of this oracle
exposed under the torch_geometric / torch_scatter module names, and commits
the resulting vectors under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path
(graph_hscn_b200/) never does: it fails loudly when libghscn.so is missing.
"""
