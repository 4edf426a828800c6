#!/usr/bin/env python3
"""bench.py — the driver-facing benchmark of the etpgt_b200 hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU oracle port, same config

Workload (BASELINE.json configs[1]): graph_transformer_optimized (D=256, L=2, H=2, k_pe=16),
BPR loss, AdamW(1e-3, 1e-5), session batches drawn from the synthetic RetailRocket-shaped graph
(etpgt_b200/synth.py).  One step = optimizer.zero_grad() + forward + loss + backward (the C++ step driver,
etpgt_gt_step_run, through etpgt_b200.train.step.FusedTrainStep) + gradient all-reduce (N > 1) + optimizer step
over one batch of `--batch` sessions per GPU (weak scaling; default 32,768; `batch_sweep` reports 16,384 and
65,536 as well).  Every step also prepares its batch (CSR + CSC index, scatter plans: etpgt_batch_prepare) inside the
timed region, one step ahead on a side stream.  Prints ONE JSON line (rank 0).

  value   sessions/s with the batch tensors already resident in HBM;
  e2e     the same step driven from pinned HOST batch tensors (H2D of x / edge_index / batch /
          targets / negatives inside the timed region) with the loss read back every step;
  roofline  the dominant kernel group (fused TransformerConv fwd+bwd) timed alone with CUDA events:
          algorithmic bytes / time vs the measured HBM peak;
  cpu_baseline  the oracle port (oracle/model_ref.py, a restatement of the reference's PyTorch/PyG
          path) on this box's host cores, on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))

METRIC = "train sessions/sec/GPU at 1/2/4/8 B200; TransformerConv edges/sec & HBM GB/s"
DIM, LAYERS, HEADS, K_PE, NUM_NEG = 256, 2, 2, 16, 5
NUM_ITEMS = 82_174


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="sessions per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=1024, help="sessions per step of the CPU sample")
    ap.add_argument("--rotate", type=int, default=4, help="distinct batches rotated through the timed steps")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--step-only", action="store_true",
                    help="only the training-step measurements (for profiler runs): no sweep, rooflines or baselines")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient / BatchNorm exchange over peer memory (this library's kernels) or NCCL all-reduces")
    ap.add_argument("--sweep-batches", type=str, default="16384,65536",
                    help="extra device-resident measurements at these batch sizes (rank 0, 1 GPU; '' = off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ helpers


def peaks() -> dict:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        d = json.loads(path.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.samples, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([f.strip() for f in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.samples:
            try:
                sm.append(float(row[0])), mx.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, row[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sample_negatives_host(rng, members, num_items, num_neg):
    """Host negatives with the reference's acceptance rule (dataloader.py:116-124): uniform in
    [1, num_items), never a session item, duplicates allowed.  Setup only (the device sampler is the
    production path once a1-a3 are on the device)."""
    out = rng.integers(1, num_items, size=(len(members), num_neg))
    for b, items in enumerate(members):
        bad = np.isin(out[b], items)
        while bad.any():
            out[b, bad] = rng.integers(1, num_items, size=int(bad.sum()))
            bad = np.isin(out[b], items)
    return out.astype(np.int64)


class HostBatch:
    """One batch as pinned host tensors, the way the reference's DataLoader hands it over."""

    FIELDS = ("x", "edge_index", "batch", "target", "negatives")

    def __init__(self, arrays: dict, pin: bool):
        self.t = {k: torch.from_numpy(np.ascontiguousarray(arrays[k])) for k in self.FIELDS}
        if pin:
            self.t = {k: v.pin_memory() for k, v in self.t.items()}
        self.num_graphs = int(arrays["target"].shape[0])
        self.nodes, self.edges = int(arrays["x"].shape[0]), int(arrays["edge_index"].shape[1])

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values())

    def to_device(self, device):
        return DeviceBatch({k: v.to(device, non_blocking=True) for k, v in self.t.items()}, self.num_graphs)


class DeviceBatch:
    def __init__(self, t: dict, num_graphs: int):
        self.x, self.edge_index, self.batch = t["x"], t["edge_index"], t["batch"]
        self.target_item, self.negative_items = t["target"], t["negatives"]
        self.num_graphs = num_graphs

    def fresh(self):
        """The same resident tensors behind a new batch object (nothing derived is carried along)."""
        return DeviceBatch({"x": self.x, "edge_index": self.edge_index, "batch": self.batch,
                            "target": self.target_item, "negatives": self.negative_items}, self.num_graphs)


def make_batches(data, edge_keys, first_session, batch, count, seed, pin):
    from etpgt_b200 import synth

    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        ids = (first_session + i * batch + np.arange(batch)) % data.num_sessions
        arrays = synth.build_batch(data, ids, edge_keys=edge_keys)
        arrays["negatives"] = sample_negatives_host(rng, arrays["members"], data.num_items, NUM_NEG)
        out.append(HostBatch(arrays, pin))
    return out


def cached_pe(num_items):
    return torch.randn(num_items, K_PE, generator=torch.Generator().manual_seed(7)).abs()


# ------------------------------------------------------------------------------ reference arm


def oracle_training_steps(data, edge_keys, batch, steps, warmup):
    """The CPU oracle port of the same step (embedding + PE, 2x TransformerConv/BN/residual, mean
    readout, BPR, AdamW over every parameter incl. the dense item table) on all host cores."""
    from oracle import model_ref

    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    state = {"item_embedding.weight": torch.randn(NUM_ITEMS, DIM, generator=g) * 0.05,
             "laplacian_pe.projection.weight": torch.randn(DIM, K_PE, generator=g) * 0.1,
             "laplacian_pe.projection.bias": torch.zeros(DIM), "laplacian_pe._cached_pe": cached_pe(NUM_ITEMS)}
    state["item_embedding.weight"][0] = 0
    for layer in range(LAYERS):
        for lin in ("query", "key", "value", "skip"):
            state[f"convs.{layer}.lin_{lin}.weight"] = torch.randn(DIM, DIM, generator=g) / DIM ** 0.5
            state[f"convs.{layer}.lin_{lin}.bias"] = torch.zeros(DIM)
        state[f"convs.{layer}.lin_beta.weight"] = torch.randn(1, 3 * DIM, generator=g) * 0.05
        state[f"batch_norms.{layer}.weight"] = torch.ones(DIM)
        state[f"batch_norms.{layer}.bias"] = torch.zeros(DIM)
        state[f"batch_norms.{layer}.running_mean"] = torch.zeros(DIM)
        state[f"batch_norms.{layer}.running_var"] = torch.ones(DIM)
    params = [k for k in state if "running" not in k and "_cached_pe" not in k]
    for k in params:
        state[k].requires_grad_(True)
    opt = torch.optim.AdamW([state[k] for k in params], lr=1e-3, weight_decay=1e-5)
    batches = make_batches(data, edge_keys, 0, batch, 2, seed=1, pin=False)
    times = []
    for step in range(warmup + steps):
        hb = batches[step % len(batches)]
        t0 = time.perf_counter()
        sess = model_ref.graph_transformer_forward(state, hb.t["x"], hb.t["edge_index"], hb.t["batch"],
                                                   num_layers=LAYERS, num_heads=HEADS, training=True)
        loss = model_ref.bpr_loss(sess, state["item_embedding.weight"], hb.t["target"], hb.t["negatives"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        loss.item()
        if step >= warmup:
            times.append(time.perf_counter() - t0)
    return batch / (sum(times) / len(times)), sum(times) / len(times)


def run_reference(args, rank):
    from etpgt_b200 import synth

    if rank != 0:
        return
    data = synth.generate()
    edge_keys = synth.sorted_edge_keys(data)
    steps, warmup = min(args.steps, 50), min(args.warmup, 5)   # 1,024-session steps: ~0.2-0.4 s each on the host cores
    value, sec = oracle_training_steps(data, edge_keys, args.cpu_batch, steps, warmup)
    cores = os.cpu_count() or 1
    sample = f"{steps} steps of {args.cpu_batch} sessions (oracle port of the reference PyTorch/PyG path, fp32, CPU)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sessions/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, data.stats()),
        "cpu_baseline": {"value": value, "unit": "sessions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sessions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, stats):
    return {"workload": "graph_transformer_optimized training step (fwd+BPR+bwd+AdamW), RR-synth sessions",
            "dim": DIM, "layers": LAYERS, "heads": HEADS, "k_pe": K_PE, "negatives": NUM_NEG,
            "sessions_per_gpu_per_step": args.batch, "items": stats["items"], "graph_edges": stats["graph_edges"],
            "graph_nodes": stats["graph_nodes"], "parallelism": f"dp{args.gpus}",
            "exchange": (args.exchange if args.gpus > 1 else "none"),
            "l2_policy": "inputs larger than L2 (rotating batches, 84 MB table, >300 MB activations)"}


# ------------------------------------------------------------------------------ B200 arm


def run_b200(args, rank, world_size, local_rank):
    import torch.distributed as dist

    import etpgt_b200
    from etpgt_b200 import _lib, ops, optim, parallel, synth
    from etpgt_b200.model import create_graph_transformer_optimized
    from etpgt_b200.train.step import FusedTrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the etpgt_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    distributed = world_size > 1
    if distributed:
        dist.init_process_group("nccl", device_id=device)

    data = synth.generate()
    edge_keys = synth.sorted_edge_keys(data)
    args.rotate = max(args.rotate, 2)   # step i + 1 is prepared while step i runs: they must be different batches
    first = rank * args.batch * args.rotate
    host_batches = make_batches(data, edge_keys, first, args.batch, args.rotate, seed=100 + rank, pin=True)
    dev_batches = [hb.to_device(device) for hb in host_batches]

    torch.manual_seed(0)
    model = create_graph_transformer_optimized(NUM_ITEMS, DIM, DIM, LAYERS, HEADS, dropout=0.1).to(device)
    model.laplacian_pe._cached_pe = cached_pe(NUM_ITEMS).to(device)
    peer = None
    if distributed:
        # before the optimizer: with the peer exchange the item table and its gradient buffer move into this rank's
        # peer region (parallel.PeerDataParallel); parameters are broadcast from rank 0
        peer = parallel.enable_data_parallel(model, exchange=args.exchange)
    opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)   # etpgt_adam_step, table-gradient sink
    params = list(model.parameters())
    total_sessions = args.batch * world_size
    model.train()

    # forward + BPR loss + backward through the C++ step driver (etpgt_gt_step_run: the same kernels as the
    # per-operator autograd path, bit-identical results, one host call); ETPGT_BENCH_AUTOGRAD=1 times that path
    fused = None if os.environ.get("ETPGT_BENCH_AUTOGRAD") else FusedTrainStep(model, "bpr")

    def step(batch):
        nonlocal total_sessions
        opt.zero_grad()
        if fused is not None:
            loss = fused(batch, total_sessions=total_sessions)[0]
            if distributed and peer is None:
                fused.allreduce_gradients()      # NCCL variant; the peer exchange is part of opt.step()
        else:
            sess = model(batch)
            loss = ops.sampled_loss(sess, model.item_embedding, batch.target_item, batch.negative_items, "bpr",
                                    total_sessions=total_sessions)[0]
            loss.backward()
            if distributed:
                parallel.allreduce_gradients(params)
        opt.step()
        return loss

    def sync():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    step_stats = {}

    def timed(fn, steps, label="value"):
        sync()
        mallocs = torch.cuda.memory_stats(device).get("num_device_alloc", 0)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        for i in range(steps):
            fn(i)
            marks[i + 1].record()
        sync()
        mallocs = torch.cuda.memory_stats(device).get("num_device_alloc", 0) - mallocs
        if mallocs and rank == 0:   # a cudaMalloc inside the timed region means the warm-up was too short
            print(f"bench.py: {mallocs} device allocations inside the timed region '{label}'", file=sys.stderr)
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        step_stats[label] = {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step),
                             "device_allocations": int(mallocs)}
        ms = torch.tensor([marks[0].elapsed_time(marks[-1])], device=device)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    losses = []
    # Batch preparation runs one step ahead on a side stream, the way the reference's DataLoader (prefetching
    # workers + pinned memory) feeds its trainer: for `e2e` the H2D copy of batch i+1, and for both `value` and
    # `e2e` its integer preparation (ops.prepare_batch: CSR + CSC index, the sorts of the two table-gradient
    # scatters — all functions of the batch's inputs only), overlap step i.  Every step still prepares (and for
    # `e2e` copies in) its own batch inside the timed region; nothing is cached from one visit of a batch to the
    # next.  The loss of step i is read back (pinned buffer + event) after step i+1 has been queued, so the
    # device never waits for the host.
    copy_stream = torch.cuda.Stream(device=device)
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    pending = {"batch": None, "ready": None, "prepared": None, "loss_event": None, "slot": 0}

    def from_host(i):
        return host_batches[i % len(host_batches)].to_device(device)

    def resident(batches):
        # same resident tensors, a fresh batch object: no index / plans carried over from its last visit
        return lambda i: batches[i % len(batches)].fresh()

    def prefetch(i, source):
        with torch.cuda.stream(copy_stream):
            batch = source(i)
            prepared = ops.prepare_batch(batch, NUM_ITEMS)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        pending["batch"], pending["ready"], pending["prepared"] = batch, ready, prepared

    def take(i, source):
        """The prepared batch of step i (made one step ahead), handed over to the compute stream; queues the
        preparation of step i + 1."""
        if pending["batch"] is None:
            prefetch(i, source)
        batch, ready, prepared = pending["batch"], pending["ready"], pending["prepared"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ready)
        for t in [batch.x, batch.edge_index, batch.batch, batch.target_item, batch.negative_items] + prepared.tensors():
            t.record_stream(cur)
        prefetch(i + 1, source)
        return batch

    def drain_loss():
        if pending["loss_event"] is not None:
            pending["loss_event"].synchronize()
            losses.append(float(loss_host[pending["slot"] ^ 1][0]))
            pending["loss_event"] = None

    e2e_marks = []

    resident_source = resident(dev_batches)

    in_flight = []

    def throttle():
        """Keeps the host at most two steps ahead of the device (a loop that logs its loss is never further
        ahead either).  Buffers handed from the preparation stream to the compute stream return to the
        allocator only when the device has passed them; an unbounded run-ahead would make it grow instead."""
        event = torch.cuda.Event()
        event.record()
        in_flight.append(event)
        if len(in_flight) > 2:
            in_flight.pop(0).synchronize()

    def value_step(i, source=None):
        step(take(i, source or resident_source))
        throttle()

    def e2e_step(i):
        if os.environ.get("ETPGT_BENCH_DEBUG"):
            e2e_marks.append(time.perf_counter())
        loss = step(take(i, from_host))
        slot = pending["slot"]
        loss_host[slot].copy_(loss.detach().reshape(1), non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        drain_loss()                      # the PREVIOUS step's loss: its copy finished long ago
        pending["loss_event"], pending["slot"] = event, slot ^ 1

    with ClockSampler(local_rank) as clocks:
        # ---- device-resident timing (value); the warm-up visits every rotating batch so that the
        # caching allocator has seen every shape before the clock starts
        for i in range(max(args.warmup, 2 * len(dev_batches) + 1)):
            value_step(i)
        pending["batch"] = None
        _lib.reset_launch_count()
        ms_total = timed(value_step, args.steps)
        launches = _lib.launch_count()
        pending["batch"] = None
        value = total_sessions * args.steps / (ms_total / 1e3)
        # ---- end to end from pinned host batches (e2e)
        # the warm-up visits every rotating batch on the copy stream too (its allocator pool must have seen
        # every shape before the clock starts, exactly as for the device-resident loop above)
        for i in range(max(args.warmup, 2 * len(host_batches) + 1)):
            e2e_step(i)
        drain_loss()
        pending["batch"] = None

        def e2e_all(i):
            e2e_step(i)
            if i == args.steps - 1:
                drain_loss()              # the last loss is read inside the timed region too

        ms_e2e = timed(e2e_all, args.steps, "e2e")
        if e2e_marks and rank == 0:
            gaps = np.diff(np.asarray(e2e_marks[-args.steps:])) * 1e3
            print("e2e host gaps between steps (ms):", np.round(gaps, 2).tolist(), file=sys.stderr)
        e2e_value = total_sessions * args.steps / (ms_e2e / 1e3)
    h2d = int(np.mean([hb.nbytes() for hb in host_batches]))

    out = {
        "metric": METRIC, "value": value, "unit": "sessions/s", "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, data.stats()),
        "e2e": {"value": e2e_value, "unit": "sessions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "step_ms": step_stats,      # device time between the ends of consecutive steps (rank 0)
        "clocks": clocks.summary(),
        "final_loss": losses[-1] if losses else None,
        "batch_shape": {"nodes": host_batches[0].nodes, "edges": host_batches[0].edges},
    }
    if args.step_only:
        if rank == 0:
            print(json.dumps(out))
        if distributed:
            dist.destroy_process_group()
        return
    if rank == 0 and world_size == 1 and args.sweep_batches:
        # The step costs the host about 1.9 ms (some 60 launches driven through Python / autograd), so small
        # batches are host-launch-bound and noisy; the sweep shows the same step on either side of the
        # default batch.
        out["batch_sweep"] = []
        for sweep in [int(v) for v in args.sweep_batches.split(",") if v]:
            if sweep == args.batch:
                continue
            big = [hb.to_device(device) for hb in make_batches(data, edge_keys, 0, sweep, 2, seed=7, pin=False)]
            total_sessions = sweep
            big_source = resident(big)
            pending["batch"] = None
            for i in range(5):
                value_step(i, big_source)
            pending["batch"] = None
            sweep_steps = max(args.steps // 2, 4)
            ms_big = timed(lambda i: value_step(i, big_source), sweep_steps, f"sweep_{sweep}")
            pending["batch"] = None
            out["batch_sweep"].append({"sessions_per_step": sweep, "ms_per_step": ms_big / sweep_steps,
                                       "value": sweep * sweep_steps / (ms_big / 1e3), "unit": "sessions/s",
                                       "nodes": big[0].x.numel(), "edges": big[0].edge_index.size(1)})
            del big
        total_sessions = args.batch * world_size
    if rank == 0:
        out["scoring"] = scoring_roofline(model, device)
        out["roofline"] = tconv_roofline(model, dev_batches[0], data, device)
        if world_size == 1:
            out["baseline_models"] = baseline_models(dev_batches, device)
        out["edges_per_s_tconv_fwd_bwd"] = out["roofline"].pop("edges_per_s")
        if world_size == 1 and not args.skip_cpu_baseline:
            steps = 60      # about 10-20 s of host work: 1,024-session steps take 0.1-0.3 s on 8-32 cores
            v, sec = oracle_training_steps(data, edge_keys, args.cpu_batch, steps, 2)
            out["cpu_baseline"] = {"value": v, "unit": "sessions/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"{steps} steps of {args.cpu_batch} sessions (oracle port, fp32 CPU)"}
        print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()


def baseline_models(dev_batches, device, steps=10):
    """BASELINE.json configs[2]: the GAT and GraphSAGE baselines (scripts/evaluate_local.py:33-58 shapes:
    3 layers, GAT with 4 averaged heads) on the same session batches: training-step sessions/s with the
    edge-softmax / mean-aggregation kernels, BPR loss and the device optimizer."""
    from etpgt_b200 import ops, optim
    from etpgt_b200.model import create_gat, create_graphsage

    out = {}
    sessions = dev_batches[0].num_graphs
    for name, make in (("gat_l3_h4", lambda: create_gat(NUM_ITEMS, DIM, DIM, 3, 4, dropout=0.1)),
                       ("graphsage_l3_mean", lambda: create_graphsage(NUM_ITEMS, DIM, DIM, 3, dropout=0.1))):
        torch.manual_seed(0)
        model = make().to(device)
        opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
        model.train()

        def step(batch):
            loss = ops.sampled_loss(model(batch), model.item_embedding, batch.target_item, batch.negative_items, "bpr")[0]
            opt.zero_grad()
            loss.backward()
            opt.step()

        # fresh batch objects: the graph index is rebuilt (inline, on the compute stream) in every step
        for i in range(4):
            step(dev_batches[i % len(dev_batches)].fresh())
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step(dev_batches[i % len(dev_batches)].fresh())
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        out[name] = {"ms_per_step": ms, "value": sessions / (ms / 1e3), "unit": "sessions/s"}
        del model, opt
    return out


def scoring_roofline(model, device, sessions=23_861, k=20, reps=10):
    """Full-catalogue evaluation scoring (BASELINE.json config 4 shape: every validation session
    against the whole item table, top-20) on the tcgen05 kernel: dense flops / time vs the measured
    bf16 tensor peak.  The item table is converted to bf16 once, as an evaluation loop does."""
    from etpgt_b200 import ops

    pk = peaks()
    table = ops.to_bf16(model.item_embedding.weight)
    sess = torch.randn(sessions, DIM, device=device) * 0.1
    sess_h = ops.to_bf16(sess)
    for _ in range(3):
        ops.score_topk(sess_h, table, k, precision="bf16")
    torch.cuda.synchronize()
    # back-to-back launches between one pair of events: the host side of a call (tensor-map encode,
    # workspace hand-out) is hidden behind the previous call's kernels, so this is device time.  The
    # working set (54 MB operands + ~1.5 GB dump buffers) is far larger than L2.
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.score_topk(sess_h, table, k, precision="bf16")
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    flops = 2.0 * sessions * NUM_ITEMS * DIM
    achieved = flops / (ms / 1e3) / 1e12
    return {"bound": "tensor", "kernel": "score_topk_tc (tcgen05 bf16 GEMM + fused top-k) + topk_merge",
            "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
            "peak_source": pk["source"], "ms": ms, "sessions": sessions, "items": NUM_ITEMS, "k": k,
            "sessions_per_s": sessions / (ms / 1e3)}


def time_tconv(qkvs, w_beta, index, device, reps=20):
    """Average CUDA-event time (ms) of the fused TransformerConv forward and backward launches on
    torch's current stream, L2 flushed between repetitions."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    n, e = index.num_nodes, index.num_edges
    f32 = dict(dtype=torch.float32, device=device)
    out, agg = torch.empty(n, DIM, **f32), torch.empty(n, DIM, **f32)
    beta, m, inv_l = torch.empty(n, **f32), torch.empty(n, HEADS, **f32), torch.empty(n, HEADS, **f32)
    d_out, d_qkvs, d_wb = torch.randn(n, DIM, **f32), torch.empty_like(qkvs), torch.empty(3 * DIM, **f32)
    ws = workspace(size("etpgt_tconv_bwd_workspace_bytes", n, e, DIM, HEADS), device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def fwd():
        call("etpgt_tconv_fwd", ptr(qkvs), n, DIM, HEADS, ptr(index.rowptr), ptr(index.col), ptr(index.eperm), e,
             ptr(w_beta), None, ptr(out), ptr(agg), ptr(beta), ptr(m), ptr(inv_l), stream())

    def bwd():
        call("etpgt_tconv_bwd", ptr(qkvs), ptr(d_out), n, DIM, HEADS, ptr(index.rowptr), ptr(index.col),
             ptr(index.eperm), ptr(index.colptr), ptr(index.row), ptr(index.cpos), e, ptr(w_beta), None, ptr(agg),
             ptr(beta), ptr(m), ptr(inv_l), ptr(d_qkvs), ptr(d_wb), ptr(ws), ws.numel(), stream())

    def time_fn(fn):
        for _ in range(3):
            fn()
        total = 0.0
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            total += a.elapsed_time(b)
        return total / reps

    return time_fn(fwd), time_fn(bwd)


def tconv_bytes(n, e):
    """Algorithmic bytes of one layer, fp32, D=256."""
    s = 4
    fwd = e * (2 * DIM * s + 4) + n * (4 * DIM * s + HEADS * 8 + 4)
    bwd = e * (4 * DIM * s + 8 + HEADS * 16 + 8) + n * (10 * DIM * s)
    return fwd, bwd


def tconv_roofline(model, batch, data, device):
    """The dominant edge-kernel group timed alone: (a) on the step's own layer-0 tensors (session batch:
    millions of tiny segments) and (b) on the whole symmetrised co-occurrence graph as ONE graph (power-law
    rows; the notebook's "whole graph as one session" case)."""
    from etpgt_b200 import ops

    pk = peaks()
    index = ops.graph_index_of(batch, batch.edge_index, batch.x.numel())
    n, e = index.num_nodes, index.num_edges
    conv = model.convs[0]
    with torch.no_grad():
        x = ops.EmbedPE.apply(batch.x, model.item_embedding.weight, model.laplacian_pe.cached(), False,
                              model.laplacian_pe.projection.weight, model.laplacian_pe.projection.bias, 0)
        qkvs = conv.project(x).contiguous()
        w_beta = conv.lin_beta.weight.detach().reshape(-1).contiguous()
    ms_f, ms_b = time_tconv(qkvs, w_beta, index, device)
    bytes_f, bytes_b = tconv_bytes(n, e)
    achieved = (bytes_f + bytes_b) / ((ms_f + ms_b) / 1e3) / 1e9
    traffic = None
    tfile = ROOT / "profiles" / "tconv_traffic.json"   # dram bytes per launch from the committed ncu capture
    if tfile.exists():
        t = json.loads(tfile.read_text())
        if t.get("nodes") == n and t.get("edges") == e:
            traffic = t["dram_bytes_fwd_bwd"]
    out = {"bound": "hbm", "kernel": "tconv_fwd + tconv_bwd_dst + tconv_bwd_src (layer 0, session batch)",
           "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
           "peak_source": pk["source"], "traffic": traffic, "ms_fwd": ms_f, "ms_bwd": ms_b,
           "gbs_fwd": bytes_f / (ms_f / 1e3) / 1e9, "gbs_bwd": bytes_b / (ms_b / 1e3) / 1e9,
           "nodes": n, "edges": e, "edges_per_s": e / ((ms_f + ms_b) / 1e3)}
    if traffic:
        # what actually crossed the HBM interface (ncu): below the algorithmic bytes because the neighbour rows
        # of a session are adjacent and are served by L1 / L2 — which is also how `frac` can exceed 1
        out["achieved_dram"] = traffic / ((ms_f + ms_b) / 1e3) / 1e9
        out["frac_dram"] = out["achieved_dram"] / pk["hbm_gbs"]
        out["note"] = ("achieved = algorithmic bytes / time; measured DRAM traffic is lower (L1/L2 reuse of "
                       "neighbour rows), see achieved_dram / frac_dram; global_graph is the same kernel group on a "
                       "working set larger than L2")
    # (b) the whole graph, both directions of every stored pair
    src = torch.from_numpy(np.concatenate([data.item_i, data.item_j])).to(device)
    dst = torch.from_numpy(np.concatenate([data.item_j, data.item_i])).to(device)
    gindex = ops.GraphIndex(torch.stack([src, dst]), data.num_items)
    with torch.no_grad():
        gq = conv.project(model.item_embedding.weight.detach()).contiguous()
    gf, gb = time_tconv(gq, w_beta, gindex, device, reps=10)
    gbf, gbb = tconv_bytes(gindex.num_nodes, gindex.num_edges)
    g_ach = (gbf + gbb) / ((gf + gb) / 1e3) / 1e9
    out["global_graph"] = {"nodes": gindex.num_nodes, "edges": gindex.num_edges, "ms_fwd": gf, "ms_bwd": gb,
                           "achieved": g_ach, "frac": g_ach / pk["hbm_gbs"],
                           "edges_per_s_fwd": gindex.num_edges / (gf / 1e3),
                           "edges_per_s_fwd_bwd": gindex.num_edges / ((gf + gb) / 1e3)}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world_size == 1 and args.gpus > 1:
        # launched without torchrun: N independent ranks are not available; measure one GPU
        print(f"bench.py: --gpus {args.gpus} without torchrun env; running a single rank", file=sys.stderr)
    run_b200(args, rank, world_size, local_rank)


if __name__ == "__main__":
    main()
