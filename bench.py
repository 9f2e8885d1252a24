#!/usr/bin/env python
"""Benchmark of the D-MPNN hot path — BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # reference CPU path (oracle port) on host cores

metric   : molecules/sec, forward + backward, D-MPNN depth 3 hidden 300 (configs[1]: batch 4096
           ZINC-size synthetic graphs per GPU, fp32, Sum read-out)
a step   : collation + CSR build of one batch -> GraphEmbedding (type ids -> [V,d],[E,d]) -> ChempropBlock -> Sum
           -> loss = H.square().mean() -> backward -> (N > 1: NCCL all-reduce of the flat gradient) -> fused Adam step
value    : whole-job molecules/s with the batch already resident in HBM (CUDA events, max over ranks)
e2e      : same step through the public API starting from pinned HOST buffers (H2D of the step's inputs
           and a D2H read of the loss inside the timed region)
roofline : dominant kernel's algorithmic bytes / its CUDA-event time / measured HBM peak
cpu_baseline : the CPU oracle port (same ATen op sequence as the reference) on this box's host cores

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "molecules/sec fwd+bwd, D-MPNN d=3 h=300"  # BASELINE.json's headline metric (configs[1]); other workloads: metric_for()
NUM_ATOM_TYPES, NUM_BOND_TYPES = 45, 13
UNIT = "molecules/s"

WORKLOADS = {
    # name: (synthetic config id, batch per GPU, hidden, depth, readout)
    "c2": dict(config=2, batch=4096, d=300, depth=3, agg="sum", desc="BASELINE configs[1]: D-MPNN depth=3 hidden=300 Sum readout, "
                                                                     "batch 4096 synthetic ZINC-size graphs per GPU, fp32 training step"),
    "c1": dict(config=1, batch=64, d=300, depth=3, agg="sum", desc="BASELINE configs[0]: batch 64 ~25-atom molecules"),
    "c3": dict(config=3, batch=16384, d=1024, depth=5, agg="mean", desc="BASELINE configs[2]: depth=5 hidden=1024 Mean readout, batch 16384 per GPU"),
    "c5": dict(config=5, batch=1024, d=2048, depth=6, agg="sum", gemm="bf16",
               desc="BASELINE configs[4]: large-molecule stress, 100-300-atom graphs, depth=6 hidden=2048, bf16 W_h with fp32 accumulation, "
                    "batch 1024 per GPU"),
    "c4": dict(config=2, batch=16384, d=300, depth=3, agg="norm", inference=True,
               desc="BASELINE configs[3]: atom message passing depth=3 hidden=300, Norm pooling, inference-only screening, 16384 molecules per launch "
                    "per GPU (10 M molecules = 611 launches; shards are independent, no collective)"),
}


def metric_for(wl: dict) -> str:
    """The metric string follows the workload: only configs[0] / configs[1] are 'D-MPNN d=3 h=300'."""
    if wl.get("inference"):
        return f"molecules/sec inference, atom message passing d={wl['depth']} h={wl['d']}, Norm read-out"
    if wl["d"] == 300 and wl["depth"] == 3:
        return METRIC
    extra = " bf16 W_h" if wl.get("gemm") == "bf16" else ""
    return f"molecules/sec fwd+bwd, D-MPNN d={wl['depth']} h={wl['d']} {wl['agg']} readout{extra}"


def workload_config(wl: dict, batch: int, V: int, E: int, n_gpus: int) -> dict:
    """The ``config`` object of the JSON line: the WORKLOAD only, so that both arms (ours / --impl reference) print the same
    object; everything about how an arm ran it goes under ``run``."""
    return {"workload": wl["desc"], "hidden": wl["d"], "depth": wl["depth"], "readout": wl["agg"], "batch_per_gpu": batch,
            "atoms_per_gpu": V, "edges_per_gpu": E, "parallelism": f"dp{n_gpus}" if not wl.get("inference") else f"replicas x{n_gpus}",
            "l2": "GPU arm: the working set of a step (> 1 GB) exceeds the 126 MB L2, no explicit flush between timed steps"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="override molecules per GPU")
    ap.add_argument("--gemm", default=None, choices=["tf32x3", "fp32", "tf32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel-table", action="store_true", help="print the per-kernel table to stderr")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one captured CUDA graph per step")
    ap.add_argument("--graph-collective", default=os.environ.get("NOTORCH_B200_GRAPH_COLLECTIVE", "on"), choices=["on", "off"],
                    help="N > 1: capture the NCCL gradient all-reduce INSIDE the step's CUDA graph (issued from autograd hooks right after layer 0's "
                         "weight gradient, overlapped with the rest of backward) instead of launching it eagerly between two graphs")
    ap.add_argument("--allreduce", default=os.environ.get("NOTORCH_B200_ALLREDUCE", "end"), choices=["overlap", "end"],
                    help="N > 1: 'end' (default) runs ONE all-reduce of the whole buffer after backward; 'overlap' issues each gradient bucket's "
                         "all-reduce from autograd hooks as soon as the bucket is final, under the rest of backward - measured 35 us SLOWER per step "
                         "at N = 8: the NCCL kernel takes SMs from persistent kernels that want all 148")
    ap.add_argument("--no-eager-cuda-baseline", action="store_true")
    ap.add_argument("--screen-molecules", type=int, default=0, help="workload c4: also time a whole screening job of this many molecules end to end")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained-clock run after the timed steps")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------

def make_workload(wl: dict, rank: int, batch: int):
    from notorch_b200.synth import config_seed, make_molecules

    seed = config_seed(wl["config"], rank)
    mols = make_molecules(batch, wl["config"], seed=seed)
    gen = torch.Generator().manual_seed(seed)
    V, E = mols.total_atoms, mols.total_edges
    # integer type features, the reference's on-the-wire input (transforms/atom.py, transforms/bond.py): 7 ids per atom
    # into a 45-entry table, 2 ids per bond into a 13-entry table
    node_types = torch.randint(0, NUM_ATOM_TYPES, (V, 7), generator=gen)
    edge_types = torch.randint(0, NUM_BOND_TYPES, (E, 2), generator=gen)
    return mols, node_types, edge_types


def algorithmic_bytes(V: int, E: int, B: int, d: int, L: int, s: int = 4) -> dict[str, float]:
    """Per-launch algorithmic bytes of each kernel class, and the whole step."""
    return {
        "K0": (V + 2 * E) * d * s + 4 * E,
        "K1": (E + V) * d * s + 4 * (E + V + 1),
        "K5": (E + V) * d * s + 4 * (E + V + 1),
        "K2": (V + 3 * E) * d * s + d * d * s + 8 * E,
        "K3": (V + B) * d * s + 4 * (B + 1),
        "K4a": 2 * E * d * s + d * d * s,  # dgrad: reads g, writes g_m
        "K4b": (V + 2 * E) * d * s + d * d * s,  # wgrad: reads g, n[src], h[rev]; writes gW
        "K6": (V + 4 * E) * d * s + 12 * E,
        "K1bwd": (V + 2 * E) * d * s + 4 * E,  # g_hL = gE + g_node[dst]
        "K3bwd": (V + B) * d * s + 4 * V,
        "K3e": (E + B) * d * s + 4 * (B + 1),  # read-out summed straight over the molecules' edge states (K1 + K3 in one pass)
        "K3ebwd": (E + B) * d * s + 4 * E,
        "Kpm": (2 * E + 2 * B) * d * s + 12 * E,  # last depth on the molecules: h and h[rev] in, M and S [B, d] out
        "K6p": 2 * (E + B) * d * s + 16 * E,  # ... and its epilogue: h in, g_h out, the [B, d] gradients from L2
        "K0e": (V * 7 + E * 2) * 8 + 4 * E + E * d * s,  # fused GraphEmbedding + edge init: type ids + src in, h0 out
        "K0ebwd": (V * 7 + E * 2) * 8 + 4 * E + E * d * s,  # one pass over g_{h0}, ids again
        "emb": 0.5 * ((V * 7 + E * 2) * 8 + (V + E) * d * s),  # two launches (atoms, bonds): average per launch
        "embbwd": 0.5 * ((V * 7 + E * 2) * 8 + (V + E) * d * s),
        "step": d * s * (6 * E + 5 * V + B + L * (12 * E + 5 * V)),
    }


def gemm_flops(E: int, d: int) -> float:
    return 2.0 * E * d * d


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        self.t_start = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            if self.t_start is not None:  # samples before mark_start() (nvidia-smi start-up, warm-up) are not of the timed region
                self.rows.append([c.strip() for c in line.split(",")])

    def mark_start(self):
        """nvidia-smi is launched BEFORE the warm-up (its NVML start-up takes the driver lock and would stall the first timed
        launches); only the samples that arrive after this call are kept."""
        self.t_start = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU leg (oracle port) — the checker used as the timed CPU baseline, nothing from it is shipped
# ------------------------------------------------------------------------------------------------

def _port_inputs(wl: dict, batch: int, nmol: int, device="cpu"):
    """Inputs of the oracle port for the first ``nmol`` molecules of the workload batch of rank 0."""
    from oracle import dmpnn_oracle as O

    mols, node_types, edge_types = make_workload(wl, 0, batch)
    sub = mols.shard(0, batch // nmol) if nmol < batch else mols
    c = O.collate(sub.split())
    V, E = sub.total_atoms, sub.total_edges
    args = (node_types[:V], edge_types[:E], torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]),
            torch.from_numpy(c["batch_node_index"]))
    return tuple(a.to(device) for a in args) + (len(sub),), sub


def _time_port(model, args, steps: int, warmup: int, sync=None) -> list[float]:
    from oracle import dmpnn_oracle as O

    for _ in range(warmup):
        O.train_step_cpu(model, *args)
    times = []
    for _ in range(steps):
        if sync:
            sync()
        t0 = time.perf_counter()
        O.train_step_cpu(model, *args)  # returns float(loss): on a CUDA model that is the device sync of the step
        times.append(time.perf_counter() - t0)
    return times


def cpu_reference_run(wl: dict, batch: int, steps: int, warmup: int, budget_s: float, threads: int | None = None) -> dict:
    from oracle import dmpnn_oracle as O

    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    # bounded sample: shrink the per-step batch until (warmup + steps) fits the budget
    sample = batch
    model = O.CpuPort(hidden_dim=wl["d"], depth=wl["depth"], agg=wl["agg"] if wl["agg"] in ("sum", "mean") else "sum",
                      embed=(NUM_ATOM_TYPES, NUM_BOND_TYPES))
    args, _ = _port_inputs(wl, batch, sample)
    t0 = time.perf_counter()
    O.train_step_cpu(model, *args)
    one = time.perf_counter() - t0
    while sample > 64 and one * (steps + warmup) > budget_s:
        sample //= 2
        args, _ = _port_inputs(wl, batch, sample)
        t0 = time.perf_counter()
        O.train_step_cpu(model, *args)
        one = time.perf_counter() - t0
    times = _time_port(model, args, steps, warmup)
    total = sum(times)
    return {"value": sample * steps / total, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} timed steps (after {warmup} warm-up) of zero_grad+forward+backward (2 EmbeddingBag + block + readout) on {sample} "
                      f"of the {batch} molecules of the workload batch, torch CPU fp32, {threads} threads",
            "ms_per_step": 1e3 * total / steps, "ms_per_step_median": 1e3 * statistics.median(times), "ms_per_step_min": 1e3 * min(times),
            "sample_molecules": sample}


def cpu_protocol_extras(wl: dict, batch: int) -> dict:
    """the one-thread number at the workload's shapes (bounded sample), BASELINE configs[0] exactly (B = 64, d = 300,
    L = 3) at one thread and at all cores with median / min over 20 iterations, and the collation of those 64 molecules timed
    separately (the oracle's restatement of BatchedGraph.from_graphs, graph.py:186-223)."""
    from oracle import dmpnn_oracle as O

    cores = os.cpu_count() or 1
    out = {}
    one = cpu_reference_run(wl, batch, steps=2, warmup=1, budget_s=8.0, threads=1)
    out["one_thread"] = {k: one[k] for k in ("value", "unit", "cores", "ms_per_step", "sample_molecules")}
    c1 = WORKLOADS["c1"]
    model = O.CpuPort(hidden_dim=c1["d"], depth=c1["depth"], agg="sum", embed=(NUM_ATOM_TYPES, NUM_BOND_TYPES))
    args, sub = _port_inputs(c1, c1["batch"], c1["batch"])
    rows = {}
    for thr in (1, cores):
        torch.set_num_threads(thr)
        t = _time_port(model, args, steps=20, warmup=5)
        rows[f"threads_{thr}"] = {"ms_per_step_median": 1e3 * statistics.median(t), "ms_per_step_min": 1e3 * min(t),
                                  "molecules_per_s": c1["batch"] / statistics.median(t)}
    parts = sub.split()
    tc = []
    for _ in range(20):
        t0 = time.perf_counter()
        O.collate(parts)
        tc.append(time.perf_counter() - t0)
    out["config1_b64"] = {"workload": c1["desc"], "atoms": sub.total_atoms, "edges": sub.total_edges, **rows,
                          "collate_ms_median": 1e3 * statistics.median(tc), "collate_ms_min": 1e3 * min(tc),
                          "protocol": "5 warm-up + 20 timed iterations of zero_grad -> forward -> H.square().mean() -> backward, time.perf_counter"}
    torch.set_num_threads(cores)
    return out


def eager_cuda_reference_run(wl: dict, batch: int, dev, steps: int = 5, warmup: int = 2, mem_budget: float = 60e9) -> dict:
    """SURVEY.md §8d / BASELINE.md §3 'same box' bar: the reference's op sequence (the oracle port = the same ATen calls: index, add,
    relu, scatter_add_ with atomics, index, sub, addmm on cuBLAS SGEMM, EmbeddingBag; autograd backward) on device='cuda', eager,
    TF32 off, same workload batch (bounded so that eager autograd's saved [E, d] tensors fit), CUDA-event timed."""
    from oracle import dmpnn_oracle as O

    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        mols, _, _ = make_workload(wl, 0, batch)
        per_mol = mols.total_edges / batch * wl["d"] * 4 * 16 * max(wl["depth"], 1)  # ~16 [E, d]-sized tensors alive per depth in eager mode
        sample = batch
        while sample > 64 and sample * per_mol > mem_budget:
            sample //= 2
        torch.manual_seed(0)
        model = O.CpuPort(hidden_dim=wl["d"], depth=wl["depth"], agg=wl["agg"] if wl["agg"] in ("sum", "mean") else "sum",
                          embed=(NUM_ATOM_TYPES, NUM_BOND_TYPES)).to(dev)
        args, _ = _port_inputs(wl, batch, sample, device=dev)
        for _ in range(warmup):
            O.train_step_cpu(model, *args)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            O.train_step_cpu(model, *args)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        return {"value": sample / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "sample_molecules": sample, "kind": "port on device='cuda'",
                "what": f"the reference's ATen op sequence, eager PyTorch on the same GPU (atomics scatter_add_, cuBLAS SGEMM, allow_tf32=False), "
                        f"{steps} timed steps after {warmup} warm-up on {sample} of the {batch} molecules; zero_grad+forward+backward, no optimizer"}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def run_reference_arm(args, wl, batch):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(wl, batch, args.steps, args.warmup, budget_s=150.0)
    mols, _, _ = make_workload(wl, 0, batch)
    line = {
        "impl": "reference", "metric": metric_for(wl), "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, batch, mols.total_atoms, mols.total_edges, args.gpus),
        "run": {"batch_per_step": res["sample_molecules"], "threads": res["cores"],
                "note": "CPU port of the reference's op sequence (oracle/dmpnn_oracle.py::CpuPort, bit-identical to the live reference per "
                        "tests/test_oracle.py); one process on the host cores whatever --gpus says"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# CUDA leg
# ------------------------------------------------------------------------------------------------

def run_screening(args, wl, batch):
    """BASELINE configs[3]: inference-only screening with the atom message-passing variant. One step = one launch of `batch`
    molecules through GraphEmbedding -> AtomMessagePassing -> Norm; ranks are independent replicas on disjoint shards (no
    collective). ``--screen-molecules M``: additionally push M molecules (M / N per rank) through the end-to-end pipeline and
    report the wall-clock time of the whole job."""
    import torch.distributed as dist

    from notorch_b200 import BatchedGraph, _lib, ops
    from notorch_b200.nn import AtomMessagePassing, GraphEmbedding, Norm

    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)  # only for the barrier / max-over-ranks of the timing
    if args.gemm:
        ops.set_gemm_mode(args.gemm)
    mols, node_types, edge_types = make_workload(wl, rank, batch)
    V, E, d, L = mols.total_atoms, mols.total_edges, wl["d"], wl["depth"]
    torch.manual_seed(0)
    embed = GraphEmbedding(NUM_ATOM_TYPES, NUM_BOND_TYPES, hidden_dim=d).to(dev).eval()
    block = AtomMessagePassing(hidden_dim=d, depth=L).to(dev).eval()
    agg = Norm(100.0)
    host = {"node_types": node_types.pin_memory(), "edge_types": edge_types.pin_memory(),
            "num_atoms": torch.from_numpy(mols.num_atoms).pin_memory(), "num_edges": torch.from_numpy(mols.num_edges).pin_memory(),
            "edge_index": torch.from_numpy(mols.edge_index).pin_memory(), "rev_index": torch.from_numpy(mols.rev_index).pin_memory()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    out_host = [torch.empty((batch, d), dtype=torch.float32).pin_memory() for _ in range(2)]

    class Packed:
        def __init__(self, t):
            self.num_atoms, self.num_edges, self.edge_index, self.rev_index = t["num_atoms"], t["num_edges"], t["edge_index"], t["rev_index"]

    def step(src):
        with torch.no_grad():
            G = BatchedGraph.from_packed(Packed(src), src["node_types"], src["edge_types"], device=dev)
            return agg(block(embed(G)))

    ops.set_index_validation("sync")
    step(resident)
    ops.set_index_validation("off")
    # three graphs: the resident batch (value leg) and two device input buffers (end-to-end leg: the H2D copy of launch t + 1 and
    # the D2H copy of launch t - 1's embeddings run on copy streams under the compute of launch t)
    devbuf = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    for bset in devbuf:
        for k, v in host.items():
            bset[k].copy_(v)
    graphs, outs = {}, {}
    launch_mode = "eager"
    if not args.no_graph:
        try:
            pool = torch.cuda.graph_pool_handle()
            for key, src in (("res", resident), (0, devbuf[0]), (1, devbuf[1])):
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        step(src)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = _lib.lib().nt_kernel_launch_count()
                with torch.cuda.graph(g, pool=pool):
                    outs[key] = step(src)
                graphs[key] = (g, _lib.lib().nt_kernel_launch_count() - n0)
            launch_mode = "cuda_graph"
        except Exception as exc:
            print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); launching eagerly", file=sys.stderr)
            graphs, outs = {}, {}
            torch.cuda.synchronize()

    def _max_over_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms

    def timed_resident(nsteps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(nsteps):
            if "res" in graphs:
                graphs["res"][0].replay()
            else:
                step(resident)
        ev1.record()
        torch.cuda.synchronize()
        return _max_over_ranks(ev0.elapsed_time(ev1))

    in_stream, out_stream = torch.cuda.Stream(), torch.cuda.Stream()

    def timed_pipeline(nsteps):
        """Host ids -> device -> embeddings -> host for `nsteps` launches, fully asynchronous: nothing waits on the host until the
        end; the wall clock and the CUDA-event time of the whole pipeline are both taken, the larger one counts."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        main = torch.cuda.current_stream()
        in_ready = [torch.cuda.Event(), torch.cuda.Event()]
        in_free = [torch.cuda.Event(), torch.cuda.Event()]     # launch t has consumed devbuf[i]
        out_ready = [torch.cuda.Event(), torch.cuda.Event()]   # launch t has written its embeddings
        out_free = [torch.cuda.Event(), torch.cuda.Event()]    # the D2H copy of launch t has drained the graph's output tensor
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        in_stream.wait_stream(main)
        out_stream.wait_stream(main)

        def prefetch(t):
            i = t & 1
            with torch.cuda.stream(in_stream):
                if t >= 2:
                    in_stream.wait_event(in_free[i])
                for k, v in host.items():
                    devbuf[i][k].copy_(v, non_blocking=True)
                in_ready[i].record(in_stream)

        prefetch(0)
        for t in range(nsteps):
            i = t & 1
            if t + 1 < nsteps:
                prefetch(t + 1)
            main.wait_event(in_ready[i])
            if t >= 2:
                main.wait_event(out_free[i])
            if i in graphs:
                graphs[i][0].replay()
                H = outs[i]
            else:
                H = step(devbuf[i])
            in_free[i].record(main)
            out_ready[i].record(main)
            with torch.cuda.stream(out_stream):
                out_stream.wait_event(out_ready[i])
                out_host[i].copy_(H, non_blocking=True)  # the screening result: one embedding per molecule
                out_free[i].record(out_stream)
            if i not in graphs:
                H.record_stream(out_stream)
        main.wait_stream(out_stream)
        ev1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        return _max_over_ranks(max(ev0.elapsed_time(ev1), wall))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize()
    if sampler:
        sampler.mark_start()
    n0 = _lib.lib().nt_kernel_launch_count()
    ms_total = timed_resident(args.steps)
    launches = graphs["res"][1] * args.steps if "res" in graphs else _lib.lib().nt_kernel_launch_count() - n0
    clocks = sampler.stop() if sampler else None
    timed_pipeline(2)
    e2e_ms = timed_pipeline(args.steps)
    screen = None
    if args.screen_molecules:
        per_rank = -(-args.screen_molecules // world)
        n_launch = -(-per_rank // batch)
        job_ms = timed_pipeline(n_launch)
        screen = {"molecules": n_launch * batch * world, "launches_per_gpu": n_launch, "seconds": job_ms * 1e-3,
                  "value": n_launch * batch * world / (job_ms * 1e-3), "unit": UNIT,
                  "what": f"{n_launch} launches of {batch} molecules on each of {world} GPU(s), host type ids + topology -> host embeddings, "
                          "wall clock of the whole job (max over ranks); the same synthetic batch is re-sent every launch (no dataset on this box)"}
    ops.set_index_validation("deferred")
    with ops.KernelTimer() as kt:
        for _ in range(min(args.steps, 10)):
            step(resident)
    summ = kt.summary()
    if rank == 0:
        nprof = min(args.steps, 10)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        s4 = d * 4
        alg = {"A1": (E + 2 * V) * s4 + 4 * (E + V + 1),  # reads ~E gathered atom rows + s_e, writes n
               "A2": 3 * V * s4 + d * s4,                   # reads n, h (residual); writes h'
               "K1": (E + V) * s4 + 4 * (E + V + 1), "K3": (V + batch) * s4 + 4 * (batch + 1),
               "emb": 0.5 * ((V * 7 + E * 2) * 8 + (V + E) * s4)}
        kernels = []
        for tag, rec in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"]):
            cls = tag.split(":")[0]
            row = {"kernel": tag, "launches_per_step": rec["launches"] / nprof, "avg_ms": rec["avg_ms"], "ms_per_step": rec["total_ms"] / nprof}
            if cls in alg:
                row["alg_bytes"] = alg[cls]
                row["gbs"] = alg[cls] / (rec["avg_ms"] * 1e-3) / 1e9
                row["hbm_frac"] = row["gbs"] / hbm_peak
            kernels.append(row)
        dom = next((k for k in kernels if "alg_bytes" in k), None)
        roof = None if dom is None else {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s",
                                         "frac": dom["hbm_frac"], "traffic": None, "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                                         "alg_bytes_per_launch": dom["alg_bytes"], "avg_launch_ms": dom["avg_ms"]}
        _emit({"metric": metric_for(wl), "value": world * batch * args.steps / (ms_total * 1e-3),
               "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(wl, batch, V, E, world),
               "run": {"gemm": ops.get_gemm_mode(), "launch": launch_mode,
                       "e2e": "two device input buffers and two pinned output buffers; H2D of launch t+1 and D2H of launch t-1 overlap the compute of launch t"},
               "clocks": clocks,
               "e2e": {"value": world * batch * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                       "d2h_bytes_per_step": batch * d * 4, "ms_per_step": e2e_ms / args.steps},
               "screening_job": screen, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": None, "kernels": kernels})
    outs.clear()
    _shutdown(world, graphs)


def allreduce_parity_check(rank: int, world: int, dev) -> dict:
    """SURVEY.md §4 item 6 on the GPUs themselves: every rank runs a small training step (its OWN 24 molecules, d = 64, L = 2, with
    GraphEmbedding) through the CUDA kernels and the bucketed NCCL all-reduce; rank 0 recomputes every rank's local gradient with the
    CPU oracle port (same seeds) and compares the exchanged buffer with their mean."""
    import torch.distributed as dist

    from notorch_b200 import BatchedGraph, ops
    from notorch_b200.nn import ChempropBlock, GraphEmbedding, Sum
    from notorch_b200.parallel import FlatGradients
    from notorch_b200.synth import make_molecules
    from oracle import dmpnn_oracle as O

    d, L, B = 64, 2, 24

    def problem(r):
        mols = make_molecules(B, 1, seed=777 + r)
        gen = torch.Generator().manual_seed(777 + r)
        return mols, torch.randint(0, NUM_ATOM_TYPES, (mols.total_atoms, 7), generator=gen), torch.randint(0, NUM_BOND_TYPES, (mols.total_edges, 2), generator=gen)

    torch.manual_seed(123)
    embed, block = GraphEmbedding(NUM_ATOM_TYPES, NUM_BOND_TYPES, hidden_dim=d), ChempropBlock(hidden_dim=d, depth=L)
    state = {"embed": {k: v.clone() for k, v in embed.state_dict().items()}, "block": {k: v.clone() for k, v in block.state_dict().items()}}
    embed, block = embed.to(dev), block.to(dev)
    flat = FlatGradients([list(block.parameters()), list(embed.parameters())], overlap=True)
    old_mode, old_gemm = ops._validate_mode, ops.get_gemm_mode()
    ops.set_index_validation("sync")
    ops.set_gemm_mode("tf32x3")  # the check is about the exchange: always in the fp32-parity mode, whatever the workload runs in
    mols, nt_, et_ = problem(rank)
    G = BatchedGraph.from_packed(mols, nt_, et_, device=dev)
    flat.zero()
    Sum()(block(embed(G))).square().mean().backward()
    flat.finish()
    ops.set_index_validation(old_mode)
    ops.set_gemm_mode(old_gemm)
    got = flat.flat.detach().cpu().double()
    out = None
    if rank == 0:
        torch.set_num_threads(os.cpu_count() or 1)
        acc = None
        for r in range(world):
            m, a, b = problem(r)
            port = O.CpuPort(hidden_dim=d, depth=L, agg="sum", embed=(NUM_ATOM_TYPES, NUM_BOND_TYPES)).double()
            with torch.no_grad():
                port.node.weight.copy_(state["embed"]["node.weight"]), port.edge.weight.copy_(state["embed"]["edge.weight"])
                for i, lin in enumerate(port.linears):
                    lin.weight.copy_(state["block"][f"layers.{i}.module.update.0.weight"])
                    lin.bias.copy_(state["block"][f"layers.{i}.module.update.0.bias"])
            c = O.collate(m.split())
            H, _, _ = port(a, b, torch.from_numpy(c["edge_index"]), torch.from_numpy(c["rev_index"]), torch.from_numpy(c["batch_node_index"]), B)
            H.square().mean().backward()
            vec = torch.cat([p.grad.reshape(-1) for lin in port.linears for p in (lin.weight, lin.bias)] + [port.node.weight.grad.reshape(-1), port.edge.weight.grad.reshape(-1)])
            acc = vec if acc is None else acc + vec
        want = acc / world
        err = float((got - want).abs().max() / want.abs().max())
        out = {"max_rel_err": err, "ranks": world, "ok": bool(err <= 1e-5), "elements": int(got.numel()),
               "what": "all-reduced CUDA flat gradient (block + embedding buckets, NCCL AVG) vs the mean of the per-rank fp64 oracle gradients; d=64 L=2 B=24 per rank"}
    dist.barrier()
    return out


def run_ours(args, wl, batch):
    import torch.distributed as dist

    from notorch_b200 import BatchedGraph, _lib, ops
    from notorch_b200.nn import ChempropBlock, GraphEmbedding, Mean, Sum
    from notorch_b200.parallel import FlatGradients

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU leg")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gemm or wl.get("gemm"):
        ops.set_gemm_mode(args.gemm or wl["gemm"])
    ops.set_index_validation("deferred")  # no per-batch device sync; an out-of-range index still raises (one batch late)

    mols, node_types, edge_types = make_workload(wl, rank, batch)
    V, E, d, L = mols.total_atoms, mols.total_edges, wl["d"], wl["depth"]
    torch.manual_seed(0)
    embed = GraphEmbedding(NUM_ATOM_TYPES, NUM_BOND_TYPES, hidden_dim=d).to(dev)
    block = ChempropBlock(hidden_dim=d, depth=L).to(dev)
    agg = (Sum if wl["agg"] == "sum" else Mean)()
    params = list(block.parameters()) + list(embed.parameters())
    use_graph = not args.no_graph
    # N > 1: two buckets. The block's gradients are final after layer 0's weight gradient (K4b), i.e. before layer 0's dgrad / backward
    # epilogue and the embedding backward: their all-reduce (1.08 MB at d = 300) is issued there from an autograd hook and runs on
    # NCCL's stream under the rest of backward; the 70 KB embedding bucket follows at the end; the mean is ReduceOp.AVG (no div kernel)
    overlap = world > 1 and args.allreduce == "overlap" and (args.graph_collective == "on" or not use_graph)
    flat = FlatGradients([list(block.parameters()), list(embed.parameters())], overlap=overlap)
    opt = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=use_graph)

    # pinned host copies (e2e leg) and device-resident copies (value leg)
    host = {"node_types": node_types.pin_memory(), "edge_types": edge_types.pin_memory(),
            "num_atoms": torch.from_numpy(mols.num_atoms).pin_memory(), "num_edges": torch.from_numpy(mols.num_edges).pin_memory(),
            "edge_index": torch.from_numpy(mols.edge_index).pin_memory(), "rev_index": torch.from_numpy(mols.rev_index).pin_memory()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    class Packed:
        def __init__(self, t):
            self.num_atoms, self.num_edges, self.edge_index, self.rev_index = t["num_atoms"], t["num_edges"], t["edge_index"], t["rev_index"]

    def fwd_bwd(src: dict, from_host: bool):
        if from_host:
            t = {k: v.to(dev, non_blocking=True) for k, v in src.items()}
        else:
            t = src
        G = BatchedGraph.from_packed(Packed(t), t["node_types"], t["edge_types"], device=dev)  # collation kernel (K-l)
        H = agg(block(embed(G)))  # embedding + CSR build + K0 + L x (K1, K2) + K1 + K3
        loss = H.square().mean()
        flat.zero()
        loss.backward()  # K3bwd, K1bwd, L x (K4b, K4a, K5, K6), K5
        return loss

    def step(src: dict, from_host: bool):
        loss = fwd_bwd(src, from_host)
        flat.finish()  # NCCL over NVLink when world > 1: joins the bucket all-reduces (overlap mode) or reduces everything now
        opt.step()
        return loss

    def capture(src: dict, from_host: bool):
        """The step captured as CUDA graphs: launched eagerly it is bound by Python (ctypes + autograd bookkeeping cost more
        than the kernels on a slow host). ONE graph per step (H2D copies, ~70 kernel launches, optimizer, D2H of the loss) - for
        several GPUs it also holds the NCCL all-reduces (--graph-collective on; captured with capture_error_mode="thread_local":
        in the default global mode the process group's watchdog thread, which polls CUDA events, invalidates or stalls the capture).
        --graph-collective off: TWO graphs (forward + backward | optimizer) with the all-reduce launched eagerly between them.
        The batch's indices are validated eagerly once (sync mode) before capture; the replayed graph skips the check."""
        ops.set_index_validation("sync")
        step(src, from_host)
        ops.set_index_validation("off")
        pinned_loss = torch.zeros(1).pin_memory()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step(src, from_host)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        g2 = None
        n0 = _lib.lib().nt_kernel_launch_count()
        # every graph of this process draws its intermediates from ONE memory pool (they are replayed one after the other and
        # nothing but the pinned loss and the parameters outlives a replay): three private pools do not fit 180 GB at configs[4]
        if world == 1 or args.graph_collective == "on":
            with torch.cuda.graph(g, pool=graph_pool, capture_error_mode="thread_local" if world > 1 else "global"):
                loss = step(src, from_host)
                pinned_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        else:
            with torch.cuda.graph(g, pool=graph_pool):
                loss = fwd_bwd(src, from_host)
                pinned_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=graph_pool):
                opt.step()
        return g, pinned_loss, _lib.lib().nt_kernel_launch_count() - n0, g2

    def replay(g: tuple):
        g[0].replay()
        if g[3] is not None:
            flat.all_reduce_mean()  # eager ReduceOp.AVG between the two graphs
            g[3].replay()

    graphs: dict[bool, tuple] = {}
    launch_mode = "eager"
    graph_pool = torch.cuda.graph_pool_handle() if use_graph else None
    if use_graph:
        try:
            graphs[False] = capture(resident, False)
            if not args.no_e2e:
                # end-to-end leg: two device input buffers; the H2D copy of step t + 1 (pinned host -> device, on a copy stream)
                # overlaps the compute of step t, the way a training loop prefetches its next batch
                devbuf = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
                for b in devbuf:
                    for k, v in host.items():
                        b[k].copy_(v)
                graphs[True] = (capture(devbuf[0], False), capture(devbuf[1], False))
            launch_mode = "cuda_graph"
        except Exception as exc:  # capture is an optimisation, never a requirement
            print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); launching eagerly", file=sys.stderr)
            graphs = {}
            devbuf = None
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            ops.set_index_validation("deferred")

    def timed(nsteps: int, src: dict, from_host: bool):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        last = None
        g = graphs.get(from_host)
        if g is not None and from_host:
            cstream = torch.cuda.Stream()
            ready = [torch.cuda.Event(), torch.cuda.Event()]

            def prefetch(i: int):
                with torch.cuda.stream(cstream):
                    for k, v in host.items():
                        devbuf[i][k].copy_(v, non_blocking=True)
                    ready[i].record(cstream)

            cstream.wait_stream(torch.cuda.current_stream())
            prefetch(0)
            for t in range(nsteps):
                i = t & 1
                torch.cuda.current_stream().wait_event(ready[i])
                replay(g[i])
                if t + 1 < nsteps:
                    prefetch(1 - i)  # its previous reader (step t - 1) has completed: we synchronised on its loss
                torch.cuda.current_stream().synchronize()
                last = float(g[i][1])  # the loss, copied device -> pinned host inside the graph
            ev1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            ms = max(ev0.elapsed_time(ev1), wall * 1e3)
            if world > 1:
                tt = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt)
            return ms, last
        for _ in range(nsteps):
            if g is not None:
                replay(g)
                continue
            loss = step(src, from_host)
            if from_host:
                last = float(loss.detach())  # D2H read of the step's result inside the timed region
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = ev0.elapsed_time(ev1)
        if from_host:
            ms = max(ms, wall * 1e3)
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, last

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step(resident, False)
    torch.cuda.synchronize()
    if sampler:
        sampler.mark_start()
    launches0 = _lib.lib().nt_kernel_launch_count()
    ms_total, _ = timed(args.steps, resident, False)
    launches = _lib.lib().nt_kernel_launch_count() - launches0
    if False in graphs:
        launches = graphs[False][2] * args.steps  # a replayed graph launches the kernels counted at capture time
    clocks = sampler.stop() if sampler else None
    value = world * batch * args.steps / (ms_total * 1e-3)

    # ---- the same step for >= 2 s back to back: the short timed region above runs at boost clocks, this one under the power cap ----
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2000.0 / max(ms_total / args.steps, 1e-3)) + 1)
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        if sampler2:
            sampler2.mark_start()
        ms_sus, _ = timed(n_sus, resident, False)
        clocks2 = sampler2.stop() if sampler2 else None
        sustained = {"value": world * batch * n_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": clocks2}

    e2e = None
    if not args.no_e2e:
        if True not in graphs:
            for _ in range(2):
                step(host, True)
        else:
            timed(2, host, True)
        e2e_ms, _ = timed(args.steps, host, True)
        e2e = {"value": world * batch * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "ms_per_step": e2e_ms / args.steps,
               "inputs": "int64 atom/bond type ids [V,7],[E,2] + packed int32 topology (counts, local edge_index, local rev_index), pinned host memory; "
                         "the copy of step t+1 overlaps the compute of step t (two device buffers, copy stream); the loss is read back every step"}

    # ---- the same step with the last depth run DENSE (DESIGN.md 5.10 switched off), same process, same batch: what the collapse is worth ----
    dense_last = None
    if world == 1 and not args.no_sustained and ops._pooled_backward and ops._fuse_readout and not wl.get("inference"):
        try:
            ops._pooled_backward = False
            if use_graph and False in graphs:
                gd = capture(resident, False)
                torch.cuda.synchronize()
                for _ in range(3):
                    gd[0].replay()
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(args.steps):
                    gd[0].replay()
                ev1.record()
                torch.cuda.synchronize()
                ms_d = ev0.elapsed_time(ev1)
                del gd
            else:
                for _ in range(3):
                    step(resident, False)
                ops._pooled_backward = False
                ms_d, _ = timed(args.steps, resident, False)
            dense_last = {"value": batch * args.steps / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d / args.steps,
                          "what": "the same step with NOTORCH_B200_POOLED_LAST=0: the last depth as K1 + K2 / K4b + K4a + K6 over the edges"}
        except Exception as exc:  # an extra figure, never a requirement
            print(f"bench.py: dense-last-depth leg failed ({type(exc).__name__}: {exc})", file=sys.stderr)
        finally:
            ops._pooled_backward = True
            ops.set_index_validation("off" if launch_mode == "cuda_graph" else "deferred")

    ops.set_index_validation("deferred")
    # ---- per-kernel CUDA-event timing (a separate instrumented pass over the same steps, launched eagerly) ----
    roof = kernels = None
    nprof = min(args.steps, 10)
    with ops.KernelTimer() as kt:  # every rank runs it (the step contains the collective); rank 0 reports
        for _ in range(nprof):
            step(resident, False)
    summ = kt.summary()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        alg = algorithmic_bytes(V, E, batch, d, L)
        kernels = []
        for tag, rec in sorted(summ.items(), key=lambda kv: -kv[1]["total_ms"]):
            cls = tag.split(":")[0]
            row = {"kernel": tag, "launches_per_step": rec["launches"] / nprof, "avg_ms": rec["avg_ms"], "ms_per_step": rec["total_ms"] / nprof}
            if cls in alg:
                row["alg_bytes"] = alg[cls]
                row["gbs"] = alg[cls] / (rec["avg_ms"] * 1e-3) / 1e9
                row["hbm_frac"] = row["gbs"] / hbm_peak
            if cls in ("K2", "K4a", "K4b"):
                row["alg_tflops"] = gemm_flops(E, d) / (rec["avg_ms"] * 1e-3) / 1e12
            kernels.append(row)
        dom = next((k for k in kernels if "alg_bytes" in k), None)
        # DRAM traffic and tensor-pipe activity per launch come from the committed ncu capture of this same step (a profiler
        # cannot run inside the timed process): profiles/r01_ncu_traffic.json, written by the round's evidence pass
        ncu = {}
        for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            try:
                ncu = json.load(open(os.path.join(ROOT, "profiles", name)))
                ncu["_source"] = f"NOT measured in this run: read from the committed capture profiles/{name} ({ncu.get('_source', 'ncu --set full')})"
                break
            except Exception:
                continue
        for k in kernels:
            rec = ncu.get(k["kernel"])
            if rec and wl is WORKLOADS["c2"] and batch == wl["batch"]:
                k["ncu_dram_bytes"] = rec["dram_bytes_per_launch"]
                k["ncu_tensor_pipe_active_pct"] = rec["tensor_pipe_active_pct"]
        if dom is not None:
            roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["hbm_frac"],
                    "traffic": dom.get("ncu_dram_bytes"), "traffic_source": ncu.get("_source") if "ncu_dram_bytes" in dom else None,
                    "tensor_pipe_active_pct": dom.get("ncu_tensor_pipe_active_pct"), "peak_source": peak_src, "alg_bytes_per_launch": dom["alg_bytes"], "avg_launch_ms": dom["avg_ms"],
                    "measured_over": f"{nprof} instrumented steps after the timed region (CUDA events around every C-ABI call)",
                    "step_alg_bytes": alg["step"], "step_hbm_frac": alg["step"] / (ms_total / args.steps * 1e-3) / 1e9 / hbm_peak}
        if args.kernel_table:
            for k in kernels:
                print(json.dumps(k), file=sys.stderr)

    # ---- N > 1: the exchanged gradient itself, on a small problem: all-reduced CUDA flat gradient == mean of the per-rank ORACLE gradients ----
    ar_check = None
    if world > 1:
        ar_check = allreduce_parity_check(rank, world, dev)

    cpu = eager = None
    if rank == 0 and world == 1:
        graphs.clear()
        devbuf = None  # noqa: F841
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        if not args.no_eager_cuda_baseline:
            try:
                eager = eager_cuda_reference_run(wl, batch, dev)
            except Exception as exc:  # e.g. out of memory at a huge workload: report, do not lose the line
                eager = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            res = cpu_reference_run(wl, batch, steps=3, warmup=1, budget_s=20.0)
            cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_step_median", "ms_per_step_min")}
            cpu.update(cpu_protocol_extras(wl, batch))

    if rank == 0:
        line = {
            "metric": metric_for(wl), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 operands, f32 accumulate / activations" if ops.get_gemm_mode() == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(wl, batch, V, E, world),
            "run": {"gemm": ops.get_gemm_mode(), "launch": launch_mode,
                    "collective": None if world == 1 else (
                        "NCCL all-reduce (AVG) of 2 gradient buckets captured in the step graph, issued from autograd hooks" if overlap and launch_mode == "cuda_graph"
                        else "ONE NCCL all-reduce (AVG) after backward, captured in the step graph" if launch_mode == "cuda_graph" and args.graph_collective == "on"
                        else "NCCL all-reduce (AVG), eager" + (" between two graphs" if launch_mode == "cuda_graph" else "")),
                    "embedding": "fused into the edge initialisation (nt_embed_edge_init)" if ops._fuse_embedding else "separate kernels",
                    "step": "collate+CSR, GraphEmbedding+edge init, ChempropBlock, readout, loss, backward, grad all-reduce (N>1), fused Adam",
                    "last_depth": ("collapsed onto the molecules (DESIGN.md 5.10): the loss reads the block through the sum read-out only, so the last depth "
                                   "computes sum_{e in b} h_L[e] and its gradients on [B, d] matrices; same outputs and gradients as the dense depth "
                                   "(tests), h_L itself is not materialised; NOTORCH_B200_POOLED_LAST=0 runs it dense")
                                  if ops._pooled_backward and ops._fuse_readout and not wl.get("inference") else "dense"},
            "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "eager_cuda_baseline": eager, "dense_last_depth": dense_last, "allreduce_check": ar_check, "kernels": kernels,
        }
        _emit(line)
    _shutdown(world, graphs)


def _shutdown(world: int, graphs: dict) -> None:
    """End of a multi-rank run. CUDA graphs that captured NCCL kernels must be destroyed BEFORE the communicator is (its teardown
    otherwise waits for them: the 2-rank run printed its line and then sat in destroy_process_group until the timeout killed it);
    and should the teardown still stall, the process leaves after 20 s - the JSON line is already on stdout."""
    import gc

    if world <= 1:
        return
    import torch.distributed as dist

    dist.barrier()
    torch.cuda.synchronize()
    graphs.clear()
    gc.collect()
    torch.cuda.synchronize()
    done = threading.Event()

    def _destroy():
        try:
            dist.destroy_process_group()
        finally:
            done.set()

    threading.Thread(target=_destroy, daemon=True).start()
    if not done.wait(20.0):
        print("bench.py: destroy_process_group did not return within 20 s; exiting", file=sys.stderr)
    sys.stderr.flush()
    os._exit(0)


def _emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) went to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to fd 1 (e.g. the NCCL version banner) must not pollute the JSON line
    args = parse_args()
    wl = WORKLOADS[args.workload]
    batch = args.batch or wl["batch"]
    if wl.get("inference") and args.impl != "reference":
        return run_screening(args, wl, batch)
    if args.impl == "reference":
        run_reference_arm(args, wl, batch)
    else:
        run_ours(args, wl, batch)


if __name__ == "__main__":
    main()
