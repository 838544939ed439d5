"""Smallest program that runs the benchmark's step a few times (target of the ncu launch list and --set full captures):
one bucket, one captured CUDA graph, 6 replays with the L2 flushed in between.   python scripts/step_once.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
j = int(sys.argv[1]) if len(sys.argv) > 1 else 0
graphs = bench.make_batches(0, 1, count=j + 1, as_graphs=True)[-1]
from graph_hscn_b200.data import Batch  # noqa: E402
step = GraphHSCNStep(StepConfig(), Batch.from_data_list(graphs), dev, padded=True, policy=BucketPolicy(444, 1024))
staged = step.make_resident(step.stage_graphs(graphs))
step.select_resident(staged)
step.capture(warmup=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(6):
    flush.zero_()
    step.select_resident(staged)
    step.run()
torch.cuda.synchronize()
print("losses", step.losses.tolist(), "nodes", staged.num_nodes, "bucket", staged.shape.n_cap, staged.shape.e_cap)
