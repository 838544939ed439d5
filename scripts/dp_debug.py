import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
def log(*a): print(f"[r{rank} {time.time()%1000:.2f}]", *a, flush=True)
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
log("init pg"); dist.init_process_group("nccl", device_id=dev); log("pg ok")
import bench
from graph_hscn_b200.train import GraphHSCNStep, StepConfig
b = bench.make_batch(rank)
step = GraphHSCNStep(StepConfig(), b, dev, padded=True); log("step built")
step.run_eager(world); torch.cuda.synchronize(); log("eager step ok", step.download().tolist())
mode = os.environ.get("CAPMODE", "global")
log("capturing, mode", mode)
import graph_hscn_b200.train as T
step.capture(world=world, warmup=2); log("captured")
for i in range(3):
    step.run(world); torch.cuda.synchronize(); log("replay", i, step.download().tolist())
dist.barrier(); torch.cuda.synchronize(); log("barrier ok")
dist.destroy_process_group(); log("done")
