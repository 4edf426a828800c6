"""Where a data-parallel training step spends its time (CUDA events on every rank, no profiler needed):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_timeline.py [peer|nccl] [rr|scaled]

Segments of one step, averaged over the timed steps, per rank (max / min over ranks printed by rank 0):
  driver      forward + loss + backward (etpgt_gt_step_run; under the peer exchange it contains the four in-kernel
              BatchNorm all-reduces, under NCCL the phase cuts + all-reduces)
  barrier_1   peer: every rank's backward is complete (waits for the slowest rank = load imbalance)
  dense_sum   peer: flat dense-gradient sum over the peers      | nccl: flat all-reduce
  table       peer: reduce-scatter + AdamW + all-gather kernel   | nccl: (rest of) the table all-reduce
  barrier_2   peer: all tables complete
  clear+adam  peer: gradient-buffer memset + dense AdamW         | nccl: replicated AdamW over every parameter
The same step on ONE rank (no exchange) is measured first as the reference point."""
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

    if peer is not None:        # split optimizer.step() of the peer exchange into its parts
        comm = peer.comm
        real_barrier, real_sum, real_update = comm.barrier, comm.sum_f32, peer.update_table
        state = {"n": 0}

        def barrier(channel=0):
            real_barrier(channel)
            state["n"] += 1
            mark("barrier_1" if state["n"] % 2 == 1 else "barrier_2")
        comm.barrier = barrier

        def sum_f32(*a):
            out = real_sum(*a)
            mark("dense_sum")
            return out
        comm.sum_f32 = sum_f32
        table_call = ops._lib.call if hasattr(ops, "_lib") else None   # noqa: F841

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
        state["n"] = 0
    if world > 1:
        dist.barrier()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    if peer is not None:
        peer.comm.check()
    names = ["start", "driver"] + (["barrier_1", "dense_sum", "barrier_2"] if peer is not None else
                                   (["allreduce"] if world > 1 else [])) + ["end"]
    seg = {}
    for a, b in zip(names[:-1], names[1:]):
        seg[f"{a}->{b}"] = float(np.mean([x.elapsed_time(y) for x, y in zip(marks[a], marks[b])]))
    seg["step"] = float(np.mean([x.elapsed_time(y) for x, y in zip(marks["start"], marks["end"])]))
    seg["step_to_step"] = float(np.mean([x.elapsed_time(y) for x, y in zip(marks["start"][:-1], marks["start"][1:])]))
    seg["nodes"] = int(np.mean([b.x.numel() for b in batches]))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, seg)
    else:
        gathered = [seg]
    if rank == 0:
        out = {"exchange": exchange if world > 1 else "none", "workload": workload, "world": world,
               "segments_ms_max_over_ranks": {k: max(g[k] for g in gathered) for k in seg},
               "segments_ms_min_over_ranks": {k: min(g[k] for g in gathered) for k in seg},
               "notes": "peer: barrier_1->dense_sum = dense sum kernel, dense_sum->barrier_2 = table kernel "
                        "(reduce-scatter + AdamW + all-gather) + barrier, barrier_2->end = gradient clear + dense AdamW; "
                        "driver->barrier_1 = wait for the slowest rank"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
