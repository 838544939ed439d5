"""Times the tcgen05 GEMM kernels alone at the bench shape (CUDA-graph replay over rotating operands).
usage: python scripts/gemm3x_bench.py [rows n k]        (also the target of the ncu --set full capture)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm  # noqa: E402


def timed(fn, iters=40, reps=5):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


def main():
    m, n, k = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (18269, 300, 300))]
    a = [torch.randn(m, k, device="cuda") for _ in range(8)]
    dy = [torch.randn(m, n, device="cuda") for _ in range(8)]
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    img = gemm.gemm3x_prep(w)
    t_nn = timed(lambda i: gemm.gemm3x(a[i % 8], img, n))
    t_tn = timed(lambda i: gemm.gemm3x_tn(dy[i % 8], a[i % 8]))
    flops = 2.0 * m * n * k
    print(f"gemm3x    [{m}x{k}] . [{n}x{k}]^T : {t_nn:7.2f} us  {flops / t_nn / 1e6:6.1f} fp32-equivalent TFLOP/s  "
          f"{(m * k + m * n) * 4 / t_nn / 1e3:6.0f} GB/s of compulsory traffic")
    print(f"gemm3x_tn [{m}x{n}]^T . [{m}x{k}] : {t_tn:7.2f} us  {flops / t_tn / 1e6:6.1f} fp32-equivalent TFLOP/s  (incl. slab reduce)")


main()
