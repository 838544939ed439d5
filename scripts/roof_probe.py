import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
batches = bench.make_batches(0, 1)
step = GraphHSCNStep(StepConfig(), batches[6], dev, padded=True, policy=BucketPolicy(444, 1024))
r = bench.spmm_roofline(step, 300); print("fresh        ", r["avg_launch_us"], r["kernel"])
staged = [step.make_resident(step.stage(b)) for b in batches]
for st in staged:
    step.select_resident(st)
    if step.graph is None:
        step.capture(warmup=0)
step.select_resident(staged[6])
r = bench.spmm_roofline(step, 300); print("after capture", r["avg_launch_us"])
for i in range(200):
    step.select_resident(staged[i % 8]); step.run()
torch.cuda.synchronize()
step.select_resident(staged[6])
r = bench.spmm_roofline(step, 300); print("after 200 steps", r["avg_launch_us"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); flush.zero_(); torch.cuda.synchronize()
r = bench.spmm_roofline(step, 300); print("after flush alloc", r["avg_launch_us"])
