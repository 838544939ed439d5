"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-step kernel shares.

usage: python profiles/summarize_launches.py gpurun_out/launches.csv [out.txt]
A step is delimited by consecutive ghscn::cast_i64_f32_kernel launches (the step casts the raw atom
features once, as its first kernel; profiles up to r1f came from a version that cast twice per step).  Times under ncu are cold-cache and serialised:
compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "cast_i64_f32" in n]
    s0, s1 = starts[-3], starts[-2]          # one full CUDA-graph replay from the timed region
    seg = rows[s0:s1]
    us = lambda r: float(r["Metric Value"].replace(",", "")) / 1000.0
    agg, cnt, tot = {}, collections.Counter(), 0.0
    for r in seg:
        n = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
        agg[n] = agg.get(n, 0.0) + us(r)
        cnt[n] += 1
        tot += us(r)
    mine = sum(v for n, v in agg.items() if "ghscn" in n)
    gemm = sum(v for n, v in agg.items() if "ghscn" not in n and re.search(r"gemm|cutlass|splitK|gemv", n, re.I))
    out = [f"source: {path}", f"kernels in one step: {len(seg)}   sum of kernel time: {tot:.1f} us (cold-cache, serialised)",
           f"hand-written ghscn kernels: {mine:.1f} us ({100 * mine / tot:.1f}%)   cuBLAS GEMM: {gemm:.1f} us "
           f"({100 * gemm / tot:.1f}%)   other torch: {tot - mine - gemm:.1f} us ({100 * (tot - mine - gemm) / tot:.1f}%)", ""]
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append(f"{v:9.1f} us {100 * v / tot:5.1f}%  x{cnt[n]:3d}  {n}")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
