"""Small-batch regime of the training step (the reference trains at 32 sessions per step, params.yaml:6): host time to
issue a step vs device time, with the step driver's launches plain or as one CUDA graph launch, and a cProfile of the
host path.

    python tools/prof_small.py [batch ...]      (on a GPU box; prints a table + the top host functions)
"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
from etpgt_b200 import data, ops, optim, synth  # noqa: E402
from etpgt_b200.model import create_graph_transformer_optimized  # noqa: E402
from etpgt_b200.train.step import FusedTrainStep  # noqa: E402

NCU = "--ncu" in sys.argv     # under ncu: a few plain (ungraphed) steps only, for the per-kernel launch list
batches = [int(v) for v in sys.argv[1:] if not v.startswith("--")] or [32, 1024]
d = synth.generate(num_sessions=40_000, graph_sessions=30_000, num_items=82_174, clusters=1600, seed=5)
graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
store = data.SessionStore(d.sess_ptr, d.sess_items)
dev = torch.device("cuda")
torch.manual_seed(0)
model = create_graph_transformer_optimized(d.num_items, 256, 256, 2, 2, dropout=0.1).to(dev)
model.laplacian_pe._cached_pe = torch.randn(d.num_items, 16, device=dev).abs()
opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
model.train()
pool = ops.BatchPreparer(depth=4)
side = torch.cuda.Stream()


def make(b, i):
    ids = (np.arange(b) + i * b) % d.num_sessions
    batch = data.build_batch(graph, store, ids)
    batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=i)
    return batch


for b in batches:
    resident = [make(b, i) for i in range(4)]
    for mode in ((False,) if NCU else (False, True)):
        fused = FusedTrainStep(model, "bpr", graph=mode)

        def step(i, prepare=True):
            batch = resident[i % 4]
            batch.__dict__.pop("_etpgt_index", None)
            prepared = None
            if prepare:
                with torch.cuda.stream(side):
                    prepared = ops.prepare_batch(batch, d.num_items, pool=pool)
                torch.cuda.current_stream().wait_stream(side)
            opt.zero_grad()
            fused(batch)
            opt.step()
            if prepared is not None:
                prepared.release()

        for i in range(4 if NCU else 20):
            step(i)
        torch.cuda.synchronize()
        if NCU:
            sys.exit(0)
        n = 300
        t0 = time.perf_counter()
        for i in range(n):
            step(i)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        # device time alone: the same steps with the host far ahead is not possible at this size, so time the
        # driver call back to back WITHOUT preparation / optimizer (events around 50 driver calls)
        ops.prepare_batch(resident[0], d.num_items)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(5):
                fused(resident[0])
            a.record()
            for _ in range(50):
                fused(resident[0])
            e.record()
        torch.cuda.synchronize()
        print(f"batch {b:6d} graph={str(mode):5s}: issue {1e3 * t_issue / n:.3f} ms/step, wall {1e3 * t_all / n:.3f} ms/step, "
              f"driver call alone {a.elapsed_time(e) / 50:.3f} ms (device-side, incl. its host issue), "
              f"graph rebuilds {fused.graph_rebuilds()}")
        model.zero_grad(set_to_none=True)
    prof = cProfile.Profile()
    prof.enable()
    for i in range(200):
        step(i)
    prof.disable()
    torch.cuda.synchronize()
    print(f"--- host profile, batch {b}, graph=True, 200 steps")
    pstats.Stats(prof).sort_stats("cumulative").print_stats(22)
