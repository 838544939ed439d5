"""How fast are the library TF32 GEMMs of the 3xTF32 path, and does cuBLASLt pick better tiles? (shapes of the bench)"""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
M, K, N = 18269, 900, 300
def graph_time(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / reps)
    return statistics.median(ts)
for lib_name in ("cublas", "cublaslt"):
    torch.backends.cuda.preferred_blas_library(lib_name)
    for tf32 in (True, False):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
        wr = torch.randn(K, N, device=dev)
        t1 = graph_time(lambda: torch.addmm(bias, a, w.t()))
        a2 = torch.randn(M, N * 3, device=dev); w2 = torch.randn(N * 3, N, device=dev)
        t2 = graph_time(lambda: torch.mm(a2, w2))
        # padded N = 320 / 384 / 512 variants of the forward
        ts = []
        for npad in (304, 320, 384):
            wp = torch.randn(npad, K, device=dev)
            ts.append(graph_time(lambda: torch.mm(a, wp.t())))
        print(f"{lib_name:9s} tf32={tf32}: fwd [M,900]x[900,300] {t1:6.1f} us   dX [M,900]x[900,300] {t2:6.1f} us   fwd N=304/320/384: "
              + " / ".join(f"{t:.1f}" for t in ts))
# plain K=300 TF32 and fp32 for reference
torch.backends.cuda.preferred_blas_library("cublas")
a = torch.randn(M, 300, device=dev); w = torch.randn(300, 300, device=dev)
for tf32 in (True, False):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    print(f"K=300 single GEMM tf32={tf32}: {graph_time(lambda: torch.mm(a, w.t())):.1f} us")
