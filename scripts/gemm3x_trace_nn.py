"""Pipeline timeline (clock stamps of CTA 0) of one ghscn_gemm3x launch at the bench shape.
Needs a trace build:  GHSCN_NVCC_EXTRA=-DGHSCN_GEMM3X_TRACE python -m graph_hscn_b200.build --force"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm
from graph_hscn_b200._lib import lib
m, n, k = 18269, 300, 300
a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda")
img = gemm.gemm3x_prep(w, False)
for _ in range(3): gemm.gemm3x(a, img, n, None, False)
torch.cuda.synchronize()
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib().call("ghscn_gemm3x_set_trace", tr.data_ptr())
gemm.gemm3x(a, img, n, None, False); torch.cuda.synchronize()
lib().call("ghscn_gemm3x_set_trace", None)
t = tr.cpu().tolist()
t0 = min(v for v in t if v)
f = lambda x: (x - t0) if x else None
print("=== gemm3x CTA 0: cycles (1 us ~ 1900)")
print("MMA g: wait_start a_full_ok b_full_ok")
for g in range(20): print(g, f(t[3*g]), f(t[3*g+1]), f(t[3*g+2]))
print("A producer of chunk g: loads_issued  empty_ok  stored  arrived")
for g in range(20): print(g, [f(t[100+4*g+i]) for i in range(4)])
print("B producer chunk g: wait_start empty_ok")
for g in range(20): print(g, f(t[300+2*g]), f(t[301+2*g]))
print("epilogue h: wait_start acc_full_ok tmem_released")
for h in range(2): print(h, f(t[400+4*h]), f(t[401+4*h]), f(t[402+4*h]))
print("end", f(t[410]))
