"""GPU check of the fused tcgen05 3xTF32 GEMM (csrc/gemm3x.cu): accuracy vs fp64, layout probes, timing.

Run on the B200 box: `timeout 300 python scripts/gemm3x_check.py [--probe] [--time]`.
Writes gpurun_out/gemm3x_check.json.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_hscn_b200 import gemm  # noqa: E402


def rel_err(c: torch.Tensor, ref: torch.Tensor) -> float:
    return float((c.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def run_case(m, n, k, transpose=False, bias=False, relu=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn((k, n) if transpose else (n, k), device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g) if bias else None
    img = gemm.gemm3x_prep(w, transpose)
    c = gemm.gemm3x(a, img, n, b, relu)
    torch.cuda.synchronize()
    wd = w.double().t() if not transpose else w.double()
    ref = a.double() @ wd
    if b is not None:
        ref = ref + b.double()
    if relu:
        ref = ref.relu()
    fp32 = a @ (w.t() if not transpose else w)
    if b is not None:
        fp32 = fp32 + b
    if relu:
        fp32 = fp32.relu()
    return {"m": m, "n": n, "k": k, "transpose": transpose, "bias": bias, "relu": relu,
            "err_gemm3x": rel_err(c, ref), "err_cublas_fp32": rel_err(fp32, ref)}


def run_tn(rows, m, n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    p = torch.randn(rows, m, device="cuda", generator=g)
    q = torch.randn(rows, n, device="cuda", generator=g)
    out = gemm.gemm3x_tn(p, q)
    torch.cuda.synchronize()
    ref = p.double().t() @ q.double()
    fp32 = p.t() @ q
    return {"tn_rows": rows, "m": m, "n": n, "err_gemm3x": rel_err(out, ref), "err_cublas_fp32": rel_err(fp32, ref)}


def probe_tn():
    """P = e_(r, mm), Q = ramp: out[mm, :] = Q[r, :]."""
    out = []
    rows, m, n = 64, 128, 32
    q = (torch.arange(rows, device="cuda").float()[:, None] * 100 + torch.arange(n, device="cuda").float()[None, :])
    for (r, mm) in [(0, 0), (1, 0), (0, 1), (0, 4), (3, 37), (9, 5), (17, 100), (40, 127)]:
        p = torch.zeros(rows, m, device="cuda")
        p[r, mm] = 1.0
        c = gemm.gemm3x_tn(p, q)
        torch.cuda.synchronize()
        nzr = sorted(set(c.nonzero()[:, 0].tolist()))
        out.append({"r": r, "mm": mm, "nonzero_rows": nzr[:8], "row": c[mm, :6].tolist(), "expect": q[r, :6].tolist(),
                    "ok": bool(torch.equal(c[mm], q[r]) and nzr in ([mm], []))})
    return out


def time_tn(rows, m, n, iters=40):
    ps = [torch.randn(rows, m, device="cuda") for _ in range(4)]
    qs = [torch.randn(rows, n, device="cuda") for _ in range(4)]
    for i in range(3):
        gemm.gemm3x_tn(ps[i], qs[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            gemm.gemm3x_tn(ps[i % 4], qs[i % 4])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return {"tn_rows": rows, "m": m, "n": n, "gemm3x_tn_us": e0.elapsed_time(e1) * 1e3 / iters}


def probe():
    """Unit-impulse probes that expose a wrong swizzle / descriptor: A = e_(r,kk) gives C[r, :] = W[:, kk]."""
    out = []
    m, n, k = 128, 32, 32
    w = (torch.arange(n, device="cuda").float()[:, None] * 100 + torch.arange(k, device="cuda").float()[None, :])
    img = gemm.gemm3x_prep(w, False)
    for (r, kk) in [(0, 0), (1, 0), (0, 1), (0, 4), (0, 8), (5, 13), (9, 31), (77, 20), (127, 7)]:
        a = torch.zeros(m, k, device="cuda")
        a[r, kk] = 1.0
        c = gemm.gemm3x(a, img, n)
        torch.cuda.synchronize()
        nz = c.nonzero()
        rows = sorted(set(nz[:, 0].tolist()))
        got = c[r].tolist() if len(rows) else []
        out.append({"r": r, "kk": kk, "nonzero_rows": rows[:8], "row_r": got[:8],
                    "expect": w[:8, kk].tolist(), "ok": bool(torch.equal(c[r], w[:, kk]) and rows in ([r], []) )})
    return out


def timeit(m, n, k, iters=50):
    a = [torch.randn(m, k, device="cuda") for _ in range(8)]       # rotate operands: 8 x 22 MB > nothing cached hot
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    img = gemm.gemm3x_prep(w)
    for i in range(5):
        gemm.gemm3x(a[i % 8], img, n)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            gemm.gemm3x(a[i % 8], img, n)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    # library baseline: split_cat + TF32 GEMM (the round-1 path)
    lin = gemm._Linear3xTF32.apply
    for i in range(3):
        lin(a[i % 8], w, None)
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for i in range(iters):
            lin(a[i % 8], w, None)
    g2.replay()
    torch.cuda.synchronize()
    e0.record()
    g2.replay()
    e1.record()
    torch.cuda.synchronize()
    us_lib = e0.elapsed_time(e1) * 1e3 / iters
    flops = 2.0 * m * n * k
    return {"m": m, "n": n, "k": k, "gemm3x_us": us, "splitcat_plus_library_us": us_lib,
            "gemm3x_eff_tflops_fp32": flops / us / 1e6, "bytes_min_mb": (m * k + m * n) * 4 / 1e6,
            "gemm3x_gbps": (m * k + m * n) * 4 / us / 1e3}


def main():
    res = {}
    if "--probe" in sys.argv:
        res["probe"] = probe()
        for p in res["probe"]:
            print(p)
    cases = [(128, 32, 32), (128, 16, 8), (300, 64, 64), (1000, 256, 256), (4096, 300, 300), (18269, 300, 300),
             (18269, 300, 300, True), (5000, 300, 300, False, True, True), (777, 304, 128), (130, 300, 44)]
    res["cases"] = []
    for cs in cases:
        r = run_case(*cs)
        res["cases"].append(r)
        print(r)
    if "--probe" in sys.argv:
        res["probe_tn"] = probe_tn()
        for p in res["probe_tn"]:
            print(p)
    for cs in [(64, 128, 32), (100, 44, 64), (1000, 300, 300), (18269, 300, 300), (18269, 256, 256), (5000, 300, 12 * 4)]:
        r = run_tn(*cs)
        res["cases"].append(r)
        print(r)
    if "--time" in sys.argv:
        res["timing_tn"] = [time_tn(18269, 300, 300), time_tn(146286, 300, 300, iters=10)]
        for t in res["timing_tn"]:
            print(t)
    if "--time" in sys.argv:
        res["timing"] = [timeit(18269, 300, 300), timeit(146286, 300, 300, iters=10), timeit(18269, 256, 256)]
        for t in res["timing"]:
            print(t)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm3x_check.json"), "w") as f:
        json.dump(res, f, indent=1)
    bad = [c for c in res["cases"] if not c["err_gemm3x"] < 6e-6]
    print("FAIL" if bad else "OK", len(bad), "bad cases")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
