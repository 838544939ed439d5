"""graph_hscn_b200 -- B200-native (sm_100a) hot path of Graph-HSCN behind the PyG operator API.

Package map:
    csrc/ + lib/libghscn.so   hand-written CUDA kernels behind the C ABI of include/ghscn.h
    _lib.py                   ctypes binding (fails loudly when the library is missing; no CPU fallback)
    ops.py                    torch custom ops `torch.ops.ghscn.*` with analytic backward
    structure.py              batch CSR / segments built once per batch, cached by tensor identity
    pyg/                      torch_geometric / torch_scatter-compatible layers and functions
    data.py, synthetic.py     batch/ptr collate convention and shape-matched synthetic batches
    models.py                 host-side mirror of the reference's MPNN / SCN / HSCN callers
    hetero.py                 on-device cluster -> virtual-node construction (K7)
    train.py                  CUDA-graph captured step, data-parallel gradient all-reduce
"""
__version__ = "0.1.0"
