"""Per-kernel times of one MinCUT pool forward + backward at B = 1024 Peptides-shaped graphs (torch.profiler).
usage: python scripts/mincut_profile.py [K] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import pyg, synthetic  # noqa: E402
from graph_hscn_b200.structure import structure_hints  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
b = synthetic.peptides_batch(1024, seed=1239)
N = b.x.size(0)
counts = b.ptr[1:] - b.ptr[:-1]
hints = dict(num_graphs=1024, batch_sorted=1, max_nodes_per_graph=int(counts.max()), no_self_loops=1)
with structure_hints(**hints):
    ei, _ = pyg.gcn_norm(b.edge_index.to(dev), None, N, add_self_loops=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, H, generator=g).to(dev).requires_grad_()
    s = torch.randn(N, K, generator=g).to(dev).requires_grad_()
    batch = b.batch.to(dev)

    def fb():
        out, adj, mc, ol = pyg.mincut_pool_ragged(x, ei, s, batch)
        (mc + ol + out.sum() * 1e-3 + adj.sum()).backward()
        s.grad = x.grad = None
    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fb()
        torch.cuda.synchronize()
print(f"K={K} H={H} N={N}")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
