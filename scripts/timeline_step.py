"""Per-stream kernel timeline of ONE CUDA-graph replay of the bench step (torch.profiler chrome trace):
start / duration / stream of every kernel, per-stream busy time, and the idle gaps of the whole device.
usage: python scripts/timeline_step.py [out.txt]"""
import json
import os
import re
import sys
import tempfile

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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        flush.zero_()
        step.run()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"]
      if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# split into replays at the big flush fills
cuts = [i for i, e in enumerate(ev) if "FillFunctor<unsigned char>" in e["name"] or (e.get("cat") == "gpu_memset" and e["dur"] > 20)]
last = ev[cuts[-1] + 1:] if cuts else ev
t0 = last[0]["ts"]
end = max(e["ts"] + e["dur"] for e in last)
lines = [f"one replay: {len(last)} device activities, span {end - t0:.1f} us"]
streams = {}
for e in last:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in streams.items():
    busy = sum(e["dur"] for e in es)
    lines.append(f"stream {s}: {len(es)} activities, busy {busy:.1f} us, first {es[0]['ts'] - t0:.1f}, last end {max(e['ts'] + e['dur'] for e in es) - t0:.1f}")
# device idle: union of intervals
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in last)
cur_s, cur_e, covered = iv[0][0], iv[0][1], 0.0
for a, b in iv[1:]:
    if a > cur_e:
        covered += cur_e - cur_s
        cur_s, cur_e = a, b
    else:
        cur_e = max(cur_e, b)
covered += cur_e - cur_s
lines.append(f"device busy (union over streams) {covered:.1f} us, idle {end - t0 - covered:.1f} us")
lines.append("")
lines.append("   start     dur  stream  name")
sid = {s: i for i, s in enumerate(streams)}
for e in last:
    name = re.sub(r"\(anonymous namespace\)::", "", e["name"])
    name = re.sub(r"\(.*", "", name)[:90]
    lines.append(f"{e['ts'] - t0:8.1f} {e['dur']:7.1f}  s{sid[e['args'].get('stream')]}  {name}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text[:6000])
