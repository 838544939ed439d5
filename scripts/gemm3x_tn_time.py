"""Times ghscn_gemm3x_tn alone (CUDA-graph replay over rotating operands)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm

rows, m, n = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (18269, 300, 300))]
iters = 20
ps = [torch.randn(rows, m, device="cuda") for _ in range(4)]
qs = [torch.randn(rows, n, device="cuda") for _ in range(4)]
for i in range(3):
    gemm.gemm3x_tn(ps[i], qs[i])
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(iters):
        gemm.gemm3x_tn(ps[i % 4], qs[i % 4])
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(5):
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
print(f"tn rows={rows} m={m} n={n}: {best:.2f} us/call (kernel + slab reduce)")
