"""Pipeline timeline (clock stamps of CTA (0,0)) of one ghscn_gemm3x_tn launch.
Needs a trace build:  GHSCN_NVCC_EXTRA=-DGHSCN_GEMM3X_TRACE python -m graph_hscn_b200.build --force"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm
from graph_hscn_b200._lib import lib
m, n, k = 18269, 300, 300
a = torch.randn(m, k, device="cuda"); dy = torch.randn(m, n, device="cuda")
for _ in range(3): gemm.gemm3x_tn(dy, a)
torch.cuda.synchronize()
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib().call("ghscn_gemm3x_set_trace", tr.data_ptr())
gemm.gemm3x_tn(dy, a); torch.cuda.synchronize()
lib().call("ghscn_gemm3x_set_trace", None)
t = tr.cpu().tolist()
t0 = min(v for v in t if v)
f = lambda x: (x - t0) if x else None
print("=== gemm3x_tn CTA(0,0): cycles; producers done", f(t[510]))
print("MMA  c: wait_start full_ok")
for c in range(48): print(c, f(t[2*c]), f(t[2*c+1]))
print("producer warp 0, its j-th chunk (c = 8 j): loads_issued->wait  empty_ok  stored  fenced  arrived")
for j in range(6): print(j, [f(t[200+5*j+i]) for i in range(5)])
print("epilogue start", f(t[500]))
