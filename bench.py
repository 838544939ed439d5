#!/usr/bin/env python
"""Benchmark of the Graph-HSCN hot path (BASELINE.json metric: fwd+bwd graphs/sec, config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one Peptides-func-shaped batch of 128 graphs per GPU:
SCN (GraphConv 9->16, ELU, Linear->K=10) + fused MinCUT losses fwd+bwd+AdamW  ->  cluster argmax and
on-device virtual-node construction  ->  HSCN (3 x HeteroConv{GAT l->v, GCN l->l, GCN v->v}, h=300,
mean readout, 2 linears) + BCE fwd+bwd+AdamW.  The timed loop is a TRAINING LOOP over 8 DIFFERENT batches
(global batches of seeds 1236..1243, train/train.py:73 iterates a DataLoader): every batch is padded into its shape
bucket and replays that bucket's CUDA graph (graph_hscn_b200/train.py).  Rank 0 prints ONE JSON line.

  value        graphs/s, whole job, the 8 packed batches resident in HBM (a 2 MB device-to-device copy into the
               bucket's static buffer is inside the timed region), CUDA-event timed per step, L2 flushed between
               timed steps, max over ranks
  e2e          same loop from pinned HOST buffers: every step's batch is copied H2D (double-buffered on a copy
               stream, overlapping the previous step) and every step's three losses are read D2H on the host
  roofline     achieved HBM GB/s of the dominant hand-written kernel (the h=300 SpMM) vs MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (oracle/step.py, the reference's algorithm) on this host's cores, bounded sample;
               its first step also checks the product's first step (same batch, same initial weights)
  dropin_eager the HSCN train loop of train/train.py:73-95 through `pyg.install()` (torch_geometric import names,
               eager layers, stock torch.optim.AdamW) over the same 8 batches
  config3      BASELINE config #3 (Peptides-struct shape, B=1024 per GPU, L1 loss, 11 targets) on the same GPUs
  --impl reference   times the CPU oracle as its own arm (PyG is not installable here)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRAPHS_PER_GPU = 128
NUM_BATCHES = 8
SEED = 1234 + 2            # manual_seed(1234 + config_id), config #2
METRIC = "Graph-HSCN fwd+bwd graphs/sec"
WORKLOAD = ("config#2: Graph-HSCN training loop over 8 distinct Peptides-func-shaped batches (128 graphs/GPU, ~151 "
            "nodes, ~307 directed edges, 9 atom feats): SCN MinCUT K=10 fwd+bwd+AdamW -> cluster argmax + virtual "
            "nodes -> HSCN 3x HeteroConv(GAT l->v, GCN l->l, GCN v->v) h=300 + BCE fwd+bwd+AdamW")


def workload_config() -> dict:
    """The `config` object both arms print (identical keys and values)."""
    return {"workload": WORKLOAD, "graphs_per_step_per_gpu": GRAPHS_PER_GPU, "batches": NUM_BATCHES,
            "batch_seeds": [SEED + j for j in range(NUM_BATCHES)], "hidden": 300, "clusters": 10, "layers": 3}


def _env_int(name: str, default: int) -> int:
    return int(os.environ.get(name, default))


def _peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed regions run (NVML)."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
def make_batches(rank: int, world: int, count: int = NUM_BATCHES, graphs_per_gpu: int = GRAPHS_PER_GPU,
                 task: str = "func", seed: int = SEED, as_graphs: bool = False):
    """`count` mini-batches of this rank.  Global batch j holds graphs_per_gpu * world graphs (seed + j); the ranks
    take EQUAL-COUNT shares balanced by node count (graphs are exchangeable inside a batch: SURVEY 8e), and every
    rank generates only its own graphs (per-graph random streams)."""
    from graph_hscn_b200 import synthetic
    from graph_hscn_b200.data import Batch
    from graph_hscn_b200.train import balanced_partition
    total = graphs_per_gpu * world
    out = []
    for j in range(count):
        sizes = [synthetic.peptides_graph_size(seed + j, i) for i in range(total)]
        mine = balanced_partition(sizes, world)[rank] if world > 1 else list(range(total))
        graphs = [synthetic.peptides_graph(seed + j, i, task) for i in mine]
        out.append(graphs if as_graphs else Batch.from_data_list(graphs))
    return out


def time_cpu_oracle(batch, steps: int, warmup: int, budget_s: float, first=None):
    """Runs oracle/step.py; shrinks the sample (graphs per step) so the run fits the time budget.  `first` (an
    OracleStep that already ran its first step, with its duration) is reused as the warm-up."""
    import torch
    from graph_hscn_b200.data import Batch
    from graph_hscn_b200.train import StepConfig
    from oracle.step import OracleStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    graphs = batch.to_data_list()
    n_graphs = len(graphs)
    if first is None:
        st = OracleStep(StepConfig(), batch)
        t0 = time.perf_counter()
        st.run()
        t_first = time.perf_counter() - t0
    else:
        st, t_first = first
    planned = (max(warmup, 1) - 1 + steps) * t_first
    if planned > budget_s:
        n_graphs = max(8, int(n_graphs * budget_s / planned))
        batch = Batch.from_data_list(graphs[:n_graphs])
        st = OracleStep(StepConfig(), batch)
        st.run()
    for _ in range(max(warmup - 1, 0)):
        st.run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        st.run()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": n_graphs * steps / total, "ms_per_step": 1e3 * total / steps, "cores": cores,
            "graphs_per_step": n_graphs, "steps": steps, "losses": st.losses}


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    batch = make_batches(0, 1, count=1)[0]
    r = time_cpu_oracle(batch, args.steps, args.warmup, budget_s=150.0)
    sample = (f"{r['graphs_per_step']} of {GRAPHS_PER_GPU} graphs per step (first batch of the loop), {r['steps']} timed "
              f"steps after {args.warmup} warm-up, torch CPU threads={r['cores']}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": r["value"], "unit": "graphs/s", "cores": r["cores"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": r["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle port of the reference path (PyG/torch_scatter are not installable offline); device cpu",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def spmm_roofline(step, hidden: int, reps: int = 40) -> dict:
    """Times the dominant hand-written kernel (GCN aggregation SpMM at width `hidden`) alone, cold L2."""
    import torch
    from graph_hscn_b200.structure import structure_cache, structure_hints
    dev = step.device
    N = step.dev["x"].size(0)
    with structure_hints(**step.hints):
        structure_cache().clear()
        step._register_blocks()
        st = structure_cache().graph(step.dev["edge_index"], N, N, False)
        w, w_t, _ = st.weights(None, normalize=True)
        d = st.by_dst
    nnz = d.num_items
    # operands larger than L2: rotate over enough (x, y) pairs that a launch never finds its operands in the
    # 126 MB L2 (10 x 2 x 23 MB = 467 MB); launched through the C ABI directly on torch's current stream
    from graph_hscn_b200._lib import lib
    from graph_hscn_b200.structure import _p, _stream
    nset = 10
    xs = [torch.randn(N, hidden, device=dev) for _ in range(nset)]
    ys = [torch.empty(N, hidden, device=dev) for _ in range(nset)]
    L, st_ = lib(), _stream()

    def launch(i):
        nonlocal st_
        L.call("ghscn_spmm", _p(d.rowptr), _p(d.col), _p(w), _p(xs[i % nset]), hidden, _p(ys[i % nset]), hidden,
               None, N, hidden, 0, st_)
    for i in range(nset):
        launch(i)
    torch.cuda.synchronize()
    # `reps` launches captured in ONE CUDA graph: the events then bracket pure device time (a Python/ctypes launch
    # costs ~5 us of CPU, more than a third of this kernel, and would otherwise be measured instead of the kernel)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st_ = _stream()
        for i in range(reps):
            launch(i)
    st_ = _stream()
    g.replay()
    torch.cuda.synchronize()
    trials = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        trials.append(a.elapsed_time(b) / reps)
    ms = statistics.mean(trials)
    algo_bytes = 4 * hidden * (N + N) + 4 * nnz + 4 * nnz + 4 * (N + 1)
    peaks = _peaks()
    achieved = algo_bytes / (ms * 1e-3) / 1e9
    traffic = traffic_note = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_note = tj.get("spmm_h300_dram_bytes_per_launch"), tj.get("how")
    return {"bound": "hbm", "kernel": f"spmm_wide_kernel<3,weighted> N={N} F={hidden} nnz={nnz}",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_note,
            "algorithmic_bytes": algo_bytes, "avg_launch_us": ms * 1e3,
            "timing": f"CUDA events around a CUDA graph of {reps} back-to-back launches (mean of 5 replays / {reps}); "
                      f"operands rotate over {nset} (x,y) sets = {2 * nset * 4 * hidden * N / 1e6:.0f} MB > L2"}


def dropin_eager(batches, dev, steps: int) -> dict:
    """train/train.py:73-95 on the drop-in route: `pyg.install()`, HSCN built from the torch_geometric.nn names,
    eager layers (no CUDA graph, no mirror-only fusions such as the two-stream schedule), stock torch.optim.AdamW.
    The hetero batches are built once (clusters from a fresh SCN, on-device K7, moved to the host like the output of
    generate_hetero_data + DataLoader collate); the timed loop moves each batch to the device every step."""
    import torch
    from graph_hscn_b200 import hetero, models, pyg
    pyg.install()
    try:
        import torch_geometric.nn as tgnn
        ns = pyg.namespace()
        assert tgnn.GCNConv is ns.GCNConv and tgnn.HeteroConv is ns.HeteroConv
        torch.manual_seed(0)
        scn = models.SCN([16], "elu", 9, 10, ops=ns).to(dev)
        host_batches = []
        for b in batches:
            bd = b.to(dev)
            with torch.no_grad():
                ei, ew = ns.gcn_norm(bd.edge_index, None, bd.x.size(0), add_self_loops=True)
                clusters = hetero.assign_clusters(torch.softmax(scn.logits(bd.x.float(), ei, ew), -1))
            host_batches.append(hetero.build_hetero_batch(bd.x, bd.edge_index, bd.batch, clusters, 10, y=bd.y).to("cpu"))
        model = models.HSCN("GAT", "GCN", "GCN", torch.relu, 9, 300, 10, 3, ops=ns).to(dev)
        first = host_batches[0].to(dev)
        model(first.x_dict, first.edge_index_dict, first)                     # lazy parameters
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=5e-4)

        def one(i):
            batch = host_batches[i % len(host_batches)].to(dev)
            pred = model(batch.x_dict, batch.edge_index_dict, batch)
            loss, _ = models.criterion("cross_entropy", pred, batch["local"].y)
            total = loss.item()
            loss.backward()
            opt.step()
            opt.zero_grad()
            return total
        for i in range(5):
            one(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            one(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": GRAPHS_PER_GPU * steps / dt, "unit": "graphs/s", "ms_per_step": 1e3 * dt / steps,
                "steps": steps, "what": "HSCN stage only (train/train.py:73-95) via pyg.install(), eager, host batches "
                                        "moved to the device every step, loss.item() every step"}
    finally:
        pyg.set_auto_device(False)
        for k in [k for k in sys.modules if k.startswith(("torch_geometric", "torch_scatter"))]:
            del sys.modules[k]


def measure_config3(dev, rank: int, world: int, steps: int = 12) -> dict:
    """BASELINE config #3: Peptides-struct shape, 1024 graphs per GPU, L1 loss on 11 targets, data-parallel."""
    import torch
    import torch.distributed as dist
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig
    cfg = StepConfig(num_classes=11, loss_fn="l1")
    batches = make_batches(rank, world, count=2, graphs_per_gpu=1024, task="struct", seed=1234 + 3)
    step = GraphHSCNStep(cfg, batches[0], dev, padded=True, policy=BucketPolicy(444, 1024, 1024, 1024))
    staged = [step.make_resident(step.stage(b)) for b in batches]
    for st in staged:
        step.select_resident(st)
        step.capture(world=world, warmup=0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for i in range(4):
        step.select_resident(staged[i % 2])
        step.run(world)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i, (a, b) in enumerate(ev):
        flush.zero_()
        a.record()
        step.select_resident(staged[i % 2])
        step.run(world)
        b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    step.release_graphs()
    ms = float(ms)
    return {"workload": "config#3: Peptides-struct shape, 1024 graphs/GPU, L1 loss, 11 targets, 2 distinct batches",
            "value": 1024 * world * steps / (ms * 1e-3), "unit": "graphs/s", "ms_per_step": ms / steps, "steps": steps,
            "n_gpus": world, "nodes_per_gpu": [s.num_nodes for s in staged]}


def run_product(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    torch.backends.cuda.matmul.allow_tf32 = False      # parity bar is 1e-5 relative: fp32 GEMMs
    torch.backends.cudnn.allow_tf32 = False
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from graph_hscn_b200._lib import lib
    from graph_hscn_b200.train import BucketPolicy, GraphHSCNStep, StepConfig

    from graph_hscn_b200.data import Batch
    cfg = StepConfig()
    graph_lists = make_batches(rank, world, as_graphs=True)
    batches = [Batch.from_data_list(graph_lists[0])]          # host-collated form of the first batch (CPU oracle legs)
    policy = BucketPolicy(max_nodes_per_graph=444, max_edges_per_graph=1024)   # Peptides dataset caps
    step = GraphHSCNStep(cfg, batches[0], dev, padded=True, policy=policy)
    # device-side collate: graphs are packed with local edge indices, `batch` / offsets are derived inside the step
    staged = [step.make_resident(step.stage_graphs(g)) for g in graph_lists]

    # ---- first-step parity check against the CPU oracle (same batch, same initial weights), 1 GPU only -----------
    parity = oracle_first = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle.step import OracleStep
        torch.set_num_threads(os.cpu_count() or 1)
        ost = OracleStep(cfg, batches[0], seed=0)
        step.scn.load_state_dict(ost.scn.state_dict())
        step.hscn.load_state_dict(ost.hscn.state_dict())
        step.select_resident(staged[0])
        step.capture(world=1, warmup=0)
        step.run(1)
        got = [float(v) for v in step.losses.cpu()]
        t0 = time.perf_counter()
        ost.run()
        oracle_first = (ost, time.perf_counter() - t0)
        err = [abs(g - w) / max(abs(w), 1e-30) for g, w in zip(got, ost.losses)]
        parity = {"first_step_losses": got, "oracle_losses": [float(v) for v in ost.losses], "rel_err": err,
                  "tolerance": 1e-5}
        if max(err) > 1e-5:
            raise SystemExit(f"bench.py: first step differs from the CPU oracle: {parity}")

    # ---- capture one graph per bucket -----------------------------------------------------------------------------
    for st in staged:
        step.select_resident(st)
        if step.graph is None:
            step.capture(world=world, warmup=0)
    # kernels of THIS library launched by one step: counted over one rolled-back eager step
    step.select_resident(staged[0])
    n0 = lib().query("ghscn_launch_count")
    step._dry_run(world, step._variant())
    launches_per_step = int(lib().query("ghscn_launch_count") - n0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(_physical_gpu_index(local_rank)) if rank == 0 else None
    nb = len(staged)
    for i in range(max(args.warmup, 3)):
        flush.zero_()
        step.select_resident(staged[i % nb])
        step.run(world)
    barrier()
    if sampler:
        sampler.start()
    # ---- device-resident timing ------------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i, (a, b) in enumerate(ev):
        flush.zero_()
        a.record()
        step.select_resident(staged[i % nb])
        step.run(world)
        b.record()
    barrier()
    per_step = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(per_step)
    # ---- end-to-end timing: pinned host batches -> H2D (double-buffered, copy stream) -> step -> D2H losses -------
    slot = step.prefetch(staged[0])
    for i in range(3):
        step.select_prefetched(staged[i % nb], slot)
        slot = step.prefetch(staged[(i + 1) % nb])
        step.run(world)
        step.download_async()
    barrier()
    t0 = time.perf_counter()
    slot = step.prefetch(staged[0])
    pending = None
    h2d = 0
    for i in range(args.steps):
        st = staged[i % nb]
        step.select_prefetched(st, slot)
        h2d += st.nbytes
        if i + 1 < args.steps:
            slot = step.prefetch(staged[(i + 1) % nb])          # overlaps this step
        step.run(world)
        nxt = step.download_async()
        if pending is not None:                                  # read the previous step's losses on the host
            pending[1].synchronize()
            losses = [float(v) for v in pending[0]]
        pending = nxt
    pending[1].synchronize()
    losses = [float(v) for v in pending[0]]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.finish() if sampler else None

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_graphs = world * GRAPHS_PER_GPU * args.steps

    roof = cpu = dropin = None
    if rank == 0:
        roof = spmm_roofline(step, cfg.hidden)
        if world == 1 and not args.no_cpu_baseline:
            r = time_cpu_oracle(batches[0], steps=3, warmup=1, budget_s=25.0, first=oracle_first)
            cpu = {"value": r["value"], "unit": "graphs/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{r['graphs_per_step']} of {GRAPHS_PER_GPU} graphs per step (first batch of the loop), 3 "
                             f"timed steps after 1 warm-up ({r['ms_per_step']:.0f} ms/step), oracle/step.py"}
        if world == 1 and not args.no_extras:
            dropin = dropin_eager([Batch.from_data_list(g) for g in graph_lists], dev, steps=40)
    config3 = None
    if not args.no_extras:
        config3 = measure_config3(dev, rank, world)
    if rank == 0:
        cfgd = workload_config()
        line = {
            "metric": METRIC, "value": total_graphs / (dev_ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd,
            "execution": {"parallelism": f"dp{world} by graph (equal-count, node-balanced shares of each global batch)",
                          "nodes_per_batch": [s.num_nodes for s in staged],
                          "buckets": sorted({(s.shape.n_cap, s.shape.e_cap) for s in staged}),
                          "cuda_graphs": step.num_graphs_captured,
                          "graph": "one CUDA graph per shape bucket (dummy-graph padding, padded virtual-node layout)",
                          "collate": "on the device (ghscn_collate_batch) from per-graph counts + local edge indices",
                          "l2": "flushed between timed steps (256 MiB write); e2e: 8 rotating batches, per-step working "
                                "set (~0.7 GB of activations) > L2, no flush",
                          "ms_per_step_min_max": [min(per_step), max(per_step)],
                          "gemm": "h x h projections: hand-written tcgen05/TMEM 3xTF32 kernels (fused hi/lo split, 3 "
                                  "accumulators), fp32-level accuracy; skinny layers hand-written"},
            "e2e": {"value": total_graphs / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": 12,
                    "how": "pinned packed batch -> H2D on a copy stream (double-buffered, overlaps the previous step) -> "
                           "D2D into the bucket's static buffer -> graph replay -> D2H of the 3 losses read on the host "
                           "one step later; wall clock over the whole loop"},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity_check": parity,
            "dropin_eager": dropin, "config3": config3, "losses_last_step": losses,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # destroy_process_group() has dead-locked while a CUDA graph that captured NCCL kernels was alive
        # (torch 2.11 / NCCL 2.28): drop the graphs first, and never let a stuck teardown hold the job
        dist.barrier()
        torch.cuda.synchronize()
        step.release_graphs()
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(20.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        try:
            dist.destroy_process_group()
        finally:
            watchdog.cancel()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ghscn", choices=["ghscn", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the dropin_eager and config3 legs")
    args = ap.parse_args()
    rank, local_rank, world = _env_int("RANK", 0), _env_int("LOCAL_RANK", 0), _env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (see module docstring)")
        run_product(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
