"""Serial vs parallel-branch HeteroConv schedule: the two must give bit-identical parameters after N steps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graph_hscn_b200.pyg import nn as pnn
from graph_hscn_b200 import train as gtrain
from graph_hscn_b200.train import GraphHSCNStep, StepConfig

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")

def run(parallel, captured, steps=12):
    pnn.PARALLEL_BRANCHES = parallel
    gtrain.TWO_STREAMS = parallel
    torch.manual_seed(0)
    step = GraphHSCNStep(StepConfig(), bench.make_batch(0), dev, padded=True)
    if captured:
        step.capture(warmup=2)
        n = steps - 2
    else:
        n = steps
    for _ in range(n):
        step.run()
    torch.cuda.synchronize()
    return step.hscn_grads.flat_param.detach().clone(), step.scn_grads.flat_param.detach().clone(), step.losses.clone()

for captured in (False, True):
    a = run(False, captured)
    b = run(True, captured)
    c = run(True, captured)
    print("captured" if captured else "eager", "hscn params equal:", torch.equal(a[0], b[0]), "scn:", torch.equal(a[1], b[1]),
          "max diff", float((a[0] - b[0]).abs().max()), "losses", a[2].tolist(), b[2].tolist(),
          "| parallel run-to-run equal:", torch.equal(b[0], c[0]))
