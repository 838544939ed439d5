"""Times ghscn_gemm3x alone (CUDA-graph replay over rotating operands). GHSCN_GEMM3X_DEBUG selects timing experiments."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm

def main():
    m, n, k = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (18269, 300, 300))]
    iters = 40
    a = [torch.randn(m, k, device="cuda") for _ in range(8)]
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    img = gemm.gemm3x_prep(w)
    for i in range(3):
        gemm.gemm3x(a[i], img, n)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            gemm.gemm3x(a[i % 8], img, n)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    print(f"dbg={os.environ.get('GHSCN_GEMM3X_DEBUG','0')} m={m} n={n} k={k}: {best:.2f} us/launch")

main()
