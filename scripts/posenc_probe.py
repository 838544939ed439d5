"""Laplacian PE precompute at the bench shape: the batched device eigensolver (csrc/posenc.cu) timed with
CUDA events, next to the reference's per-graph host loop (oracle/posenc.py: get_laplacian -> np.linalg.eigh float32 ->
get_lap_decomp_stats) on a bounded sample.  PROBE_GRAPHS picks the batch.  python scripts/posenc_probe.py"""
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import posenc, synthetic  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("PROBE_GRAPHS", "128"))
b = synthetic.peptides_batch(B, seed=1236)
N = b.x.size(0)
counts = b.ptr[1:] - b.ptr[:-1]
ei = b.edge_index.to(dev)
vals, vecs, sweeps = posenc.laplacian_eig(ei, b.ptr, N, int(counts.max()), return_sweeps=True)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    posenc.laplacian_eig(ei, b.ptr, N, int(counts.max()))
    e.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(e))
ms = statistics.median(ts)
print(f"B={B} N={N} max_n={int(counts.max())} sweeps {int(sweeps.min())}-{int(sweeps.max())}: "
      f"{ms:.2f} ms per batch = {B / ms * 1e3:.0f} graphs/s")
from oracle import posenc as op  # noqa: E402  (CPU baseline leg only)
sample = min(B, 32)
t0 = time.perf_counter()
for gi in range(sample):
    lo, hi = int(b.ptr[gi]), int(b.ptr[gi + 1])
    m = (b.edge_index[0] >= lo) & (b.edge_index[0] < hi)
    op.compute_posenc_stats(b.edge_index[:, m] - lo, hi - lo, True)
dt = time.perf_counter() - t0
print(f"reference host loop (oracle port, {torch.get_num_threads()} threads, {sample} graphs): "
      f"{sample / dt:.0f} graphs/s")
