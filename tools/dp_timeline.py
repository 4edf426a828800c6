"""Where a data-parallel training step spends its time (CUDA events on every rank, no profiler needed):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_timeline.py [peer|nccl] [rr|scaled]

Prints, per mark, the mean offset from the start of its step (max / min over ranks): `driver` = end of forward + loss +
backward; peer exchange: `table_grad_ready`, `side_barrier_in`, `side_table_done` (the table's reduce-scatter + AdamW +
all-gather on the exchange stream), `main_barrier` (every rank's backward complete: waits for the slowest rank),
`dense_sum`, `end`; NCCL: `allreduce`, `end`.  Run it with one process for the single-GPU reference point."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from etpgt_b200 import data as ddata, ops, optim, parallel  # noqa: E402
from etpgt_b200.model import create_graph_transformer_optimized  # noqa: E402
from etpgt_b200.train.step import FusedTrainStep  # noqa: E402


def main():
    exchange = sys.argv[1] if len(sys.argv) > 1 else "peer"
    workload = sys.argv[2] if len(sys.argv) > 2 else "rr"
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    synth = bench.load_synth()
    num_items = bench.WORKLOADS[workload]["items"]
    batch_size, rotate, steps = 32768, 4, 20
    if workload == "scaled":
        d = synth.generate_scaled(build_graph=False)
        gi, gj, _, _ = ddata.build_co_event_graph(d.sess_ptr, d.sess_items, None, 5, num_items, dev)
        graph = ddata.ItemGraph(gi, gj, num_items, dev)
    else:
        d = synth.generate()
        graph = ddata.ItemGraph(d.item_i, d.item_j, num_items, dev)
    store = ddata.SessionStore(d.sess_ptr, d.sess_items, dev)
    batches = []
    for i in range(rotate):
        ids = torch.from_numpy((rank * batch_size * rotate + i * batch_size + np.arange(batch_size)) % d.num_sessions).to(dev)
        b = ddata.build_batch(graph, store, ids, 50, False, False)
        b.negative_items = ddata.sample_negatives(store, ids, num_items, 5, seed=3, step=i)
        ops.prepare_batch(b, num_items)
        batches.append(b)
    torch.manual_seed(0)
    model = create_graph_transformer_optimized(num_items, 256, 256, 2, 2, dropout=0.1).to(dev)
    model.laplacian_pe._cached_pe = bench.cached_pe(num_items).to(dev)
    peer = parallel.enable_data_parallel(model, exchange=exchange) if world > 1 else None
    opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    fused = FusedTrainStep(model, "bpr")
    model.train()
    total = batch_size * world
    marks = {}

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.setdefault(name, []).append(e)

    if peer is not None:        # marks inside optimizer.step() of the peer exchange (each on the stream it runs on)
        comm = peer.comm
        real_barrier, real_sum, real_ready = comm.barrier, comm.sum_f32, peer.mark_table_ready
        state = {"side": 0}

        def barrier(channel=0):
            real_barrier(channel)
            if channel == 0:
                mark("main_barrier")
            else:
                state["side"] += 1
                mark("side_barrier_in" if state["side"] % 2 == 1 else "side_table_done")
        comm.barrier = barrier

        def sum_f32(*a):
            out = real_sum(*a)
            mark("dense_sum")
            return out
        comm.sum_f32 = sum_f32

        def ready():
            real_ready()
            mark("table_grad_ready")
        peer.mark_table_ready = ready

    def step(i):
        b = batches[i % rotate]
        opt.zero_grad()
        mark("start")
        fused(b, total_sessions=total)
        mark("driver")
        if world > 1 and peer is None:
            fused.allreduce_gradients()
            mark("allreduce")
        opt.step()
        mark("end")

    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    marks.clear()
    if peer is not None:
        state["side"] = 0
    if world > 1:
        dist.barrier()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    if peer is not None:
        peer.comm.check()
    # a timeline: mean offset (ms) of every mark from the start of its step
    seg = {}
    for name, events in marks.items():
        if name != "start" and len(events) == len(marks["start"]):
            seg[name] = float(np.mean([x.elapsed_time(y) for x, y in zip(marks["start"], events)]))
    seg["step_to_step"] = float(np.mean([x.elapsed_time(y) for x, y in zip(marks["start"][:-1], marks["start"][1:])]))
    seg["nodes"] = int(np.mean([b.x.numel() for b in batches]))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, seg)
    else:
        gathered = [seg]
    if rank == 0:
        out = {"exchange": exchange if world > 1 else "none", "workload": workload, "world": world,
               "offsets_ms_max_over_ranks": dict(sorted(((k, max(g[k] for g in gathered)) for k in seg), key=lambda kv: kv[1])),
               "offsets_ms_min_over_ranks": dict(sorted(((k, min(g[k] for g in gathered)) for k in seg), key=lambda kv: kv[1])),
               "notes": "mean offset of each mark from the start of its step; peer exchange: table_grad_ready (end of the "
                        "phase that completes the table gradient) -> side_barrier_in -> side_table_done is the table's "
                        "reduce-scatter + AdamW + all-gather on the exchange stream, underneath driver (end of backward) "
                        "-> main_barrier (wait for the slowest rank) -> dense_sum -> end (dense AdamW, join)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
