"""Real (warm, back-to-back) per-kernel times of one CUDA-graph replay of the bench step via torch.profiler (CUPTI).
usage: python scripts/profile_step.py [out.txt]   -- complements the cold-cache ncu launch lists in profiles/."""
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
batch = bench.make_batches(0, 1, count=int(os.environ.get("PROFILE_BATCH", "0")) + 1)[-1]
step = GraphHSCNStep(StepConfig(), batch, dev, padded=True, policy=BucketPolicy(444, 1024))
step.capture(warmup=3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
REPS = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(REPS):
        flush.zero_()
        step.run()
    torch.cuda.synchronize()
agg, cnt = collections.defaultdict(float), collections.Counter()
first, last = None, None
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(anonymous namespace\)::", "", e.name)
        name = re.sub(r"\(.*", "", name)[:95]
        if "FillFunctor<unsigned char>" in e.name or "Memset" in e.name and e.device_time > 30:
            continue
        agg[name] += e.device_time / REPS
        cnt[name] += 1
tot = sum(agg.values())
mine = sum(v for n, v in agg.items() if "ghscn" in n)
gemm = sum(v for n, v in agg.items() if "ghscn" not in n and re.search(r"gemm|cutlass|splitK|gemv", n, re.I))
lines = [f"torch.profiler (CUPTI), {REPS} graph replays, L2 flushed between replays; per-replay averages",
         f"sum of kernel time per step: {tot:.1f} us   ghscn: {mine:.1f} us ({100*mine/tot:.1f}%)   "
         f"library GEMM: {gemm:.1f} us ({100*gemm/tot:.1f}%)   other torch: {tot-mine-gemm:.1f} us", ""]
for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:60]:
    lines.append(f"{v:9.1f} us {100*v/tot:5.1f}%  x{cnt[n]//REPS:3d}  {n}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
