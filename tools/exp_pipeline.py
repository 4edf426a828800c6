"""Experiment: where batch preparation should run.  Times the bench step (32,768 sessions) with
 (a) cached index, inline scatter sorts (the pre-plan behaviour);  (b) prepare_batch inline on the compute
 stream every step;  (c) prepare_batch one step ahead on a side stream, host throttled to k steps ahead."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from etpgt_b200 import ops, optim, synth  # noqa: E402
from etpgt_b200.model import create_graph_transformer_optimized  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
data = synth.generate()
keys = synth.sorted_edge_keys(data)
dev = torch.device("cuda")
hb = bench.make_batches(data, keys, 0, B, 4, seed=1, pin=True)
db = [h.to_device(dev) for h in hb]
torch.manual_seed(0)
model = create_graph_transformer_optimized(bench.NUM_ITEMS, 256, 256, 2, 2, dropout=0.1).to(dev)
model.laplacian_pe._cached_pe = bench.cached_pe(bench.NUM_ITEMS).to(dev)
opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
model.train()
side = torch.cuda.Stream()


def step(b):
    sess = model(b)
    loss = ops.sampled_loss(sess, model.item_embedding, b.target_item, b.negative_items, "bpr", total_sessions=B)[0]
    opt.zero_grad()
    loss.backward()
    opt.step()


from etpgt_b200.train.step import FusedTrainStep  # noqa: E402

fused = FusedTrainStep(model, "bpr")


def step_driver(b):
    opt.zero_grad()
    fused(b, total_sessions=B)
    opt.step()


def run(name, fn, steps=20, reset=None):
    for rep in range(3):
        if reset:
            reset()
        for i in range(6):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        host = (time.perf_counter() - t0) / steps * 1e3
        torch.cuda.synchronize()
        print(f"{name:40s} {a.elapsed_time(b) / steps:.3f} ms/step (host enqueue {host:.3f} ms/step)", flush=True)


run("a cached index, inline sorts", lambda i: step(db[i % 4]))
run("b prepare inline (compute stream)", lambda i: (lambda bt: (ops.prepare_batch(bt, bench.NUM_ITEMS), step(bt)))(db[i % 4].fresh()))
run("b' index inline, no plans", lambda i: step(db[i % 4].fresh()))

run("d driver, cached index, inline sorts", lambda i: step_driver(db[i % 4]))
state = {"p": None, "ev": []}


def piped(k, record=True):
    def fn(i):
        if state["p"] is None:
            with torch.cuda.stream(side):
                bt = db[i % 4].fresh()
                pr = ops.prepare_batch(bt, bench.NUM_ITEMS)
                ev = torch.cuda.Event(); ev.record(side)
            state["p"] = (bt, pr, ev)
        bt, pr, ev = state["p"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        if record:
            for t in pr.tensors():
                t.record_stream(cur)
        with torch.cuda.stream(side):
            nb = db[(i + 1) % 4].fresh()
            npr = ops.prepare_batch(nb, bench.NUM_ITEMS)
            nev = torch.cuda.Event(); nev.record(side)
        state["p"] = (nb, npr, nev)
        (step_driver if DRIVER else step)(bt)
        if k:
            e = torch.cuda.Event(); e.record(); state["ev"].append(e)
            if len(state["ev"]) > k:
                state["ev"].pop(0).synchronize()
    return fn


def reset():
    state["p"], state["ev"] = None, []


DRIVER = False
for k in (1, 2):
    run(f"c side stream, host <= {k} ahead", piped(k), reset=reset)
DRIVER = True
for k in (1, 2):
    run(f"e driver + side stream, host <= {k} ahead", piped(k), reset=reset)
run("a again", lambda i: step(db[i % 4]))
run("d again", lambda i: step_driver(db[i % 4]))

if len(sys.argv) > 2:
    import cProfile
    import pstats
    reset()
    fn = lambda i: step_driver(db[i % 4])
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        fn(i)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumtime").print_stats(45)
