"""Data-parallel parity check on real GPUs (run under torchrun, any world size >= 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py [peer|nccl]

`peer` (default): BatchNorm sums, dense gradients and the table's reduce-scatter + AdamW + all-gather run as this
library's kernels over peer memory (parallel.PeerDataParallel); `nccl`: torch.distributed all-reduces between the
phases of the step driver and a replicated optimizer.

Every rank trains on its contiguous share of the same global session batches (global BatchNorm
statistics, loss divided by the global batch, gradient all-reduce, device optimizer with the
table-gradient sink); rank 0 also trains a single-process replica on the WHOLE batches.  After a few
steps the parameters must agree to fp32 rounding, and the item-sharded top-k must equal the
single-GPU top-k exactly.
"""

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))

from etpgt_b200 import data, ops, optim, parallel, synth  # noqa: E402
from etpgt_b200.model import create_graph_transformer_optimized  # noqa: E402
from etpgt_b200.train.step import FusedTrainStep  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    exchange = sys.argv[1] if len(sys.argv) > 1 else "peer"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    d = synth.generate(num_sessions=6000, graph_sessions=4500, num_items=20000, clusters=200, seed=5)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    global_batch, steps = 1024, 3

    def make(dp):
        torch.manual_seed(1)
        model = create_graph_transformer_optimized(d.num_items, 256, 256, dropout=0.0).cuda()
        model.laplacian_pe._cached_pe = torch.randn(d.num_items, 16, generator=torch.Generator().manual_seed(7)).abs().cuda()
        if dp:
            peer = parallel.enable_data_parallel(model, exchange=exchange)     # before the optimizer
        else:
            model.bn_process_group = False
        return model, optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)

    def train(model, opt, dp):
        """The data-parallel replica trains through the C++ step driver (phase by phase, BatchNorm sums
        all-reduced in between, flat gradient all-reduce); the single-process reference through the
        per-operator autograd path."""
        model.train()
        first_grads = None
        fused = FusedTrainStep(model, "bpr") if dp else None
        for step in range(steps):
            ids = np.arange(step * global_batch, (step + 1) * global_batch)
            if dp:
                cost = (d.sess_ptr[ids + 1] - d.sess_ptr[ids]).astype(np.float64)
                cuts = parallel.partition_sessions(cost, world)
                ids = ids[cuts[rank]:cuts[rank + 1]]
            batch = data.build_batch(graph, store, ids, 50, False, False)
            neg = data.sample_negatives(store, ids, d.num_items, 5, seed=3, step=step)
            opt.zero_grad()
            if dp:
                batch.negative_items = neg
                ops.prepare_batch(batch, d.num_items)
                loss = fused(batch, batch.target_item, neg, total_sessions=global_batch)[0]
                fused.allreduce_gradients()
            else:
                loss = ops.sampled_loss(model(batch), model.item_embedding, batch.target_item, neg, "bpr",
                                        total_sessions=global_batch)[0]
                loss.backward()
            if step == 0:
                first_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
                peer = getattr(model, "_etpgt_peer", None)
                if dp and peer is not None:    # the summed table gradient is never formed on the training path
                    first_grads["item_embedding.weight"] = peer.reduced_table_gradient()
            opt.step()
        if dp and getattr(model, "_etpgt_peer", None) is not None:
            model._etpgt_peer.comm.check()
        return model, first_grads, loss.item()

    def eval_sessions(model):
        model.eval()
        with torch.no_grad():
            out = model(data.build_batch(graph, store, np.arange(5000, 5256), 50, False, False))
        model.train()
        return out

    dp_model, dp_grads, dp_loss = train(*make(world > 1), world > 1)
    ok = True
    if world > 1:   # every rank must hold bit-identical parameters after training
        sums = torch.stack([p.detach().double().sum() for p in dp_model.parameters()])
        gathered = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        if not all(torch.equal(gathered[0], g) for g in gathered):
            ok = False
            if rank == 0:
                print("MISMATCH: parameters differ between ranks", (gathered[0] - gathered[1]).abs().max().item())
    if rank == 0:
        ref, ref_grads, ref_loss = train(*make(False), False)
        # (a) the global-batch gradient of the first step: all-reduced DP gradient == single-process gradient
        scale = max(g.double().abs().max().item() for g in ref_grads.values())
        worst = 0.0
        for name, g in ref_grads.items():
            err = (dp_grads[name].double() - g.double()).abs().max().item() / scale
            worst = max(worst, err)
            if err > 1e-5:
                ok = False
                print(f"MISMATCH grad {name}: {err:.3e}")
        print(f"exchange = {exchange}")
        print(f"dp{world} vs single process: step-0 gradient, worst difference / largest gradient = {worst:.3e}")
        # (b) after `steps` optimizer steps: outputs agree (raw weights are not compared: Adam turns
        # rounding-level gradient elements into +-lr steps of arbitrary sign on either path)
        a, b = eval_sessions(dp_model), eval_sessions(ref)
        err = (a.double() - b.double()).abs().max().item() / b.double().abs().max().item()
        print(f"after {steps} steps: eval session embeddings differ by {err:.3e} (relative), "
              f"last losses {dp_loss:.6f} (this rank's share) / {ref_loss:.6f}")
        ok = ok and err < 5e-3
    # item-sharded evaluation: identical ids to the single-GPU scorer
    dp_model.eval()
    ids = np.arange(4500, 4500 + 512)
    cuts = parallel.partition_sessions(np.ones(len(ids)), world)
    with torch.no_grad():
        mine = data.build_batch(graph, store, ids[cuts[rank]:cuts[rank + 1]], 50, False, False)
        sess = dp_model(mine)
        same = True
        for precision in ("fp32", "bf16"):     # exact for either scorer
            top = parallel.sharded_predict(dp_model, sess, k=20, precision=precision)
            single = ops.score_topk(sess, dp_model.get_item_embeddings(), 20, precision=precision)[1]
            same = same and bool(torch.equal(top, single))
    if world > 1:
        flag = torch.tensor([int(same), int(ok)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same, ok = bool(flag[0].item()), bool(flag[1].item())
    if rank == 0:
        print(f"sharded top-20 == single-GPU top-20 on every rank: {same}")
        print("DP CHECK", "PASSED" if (ok and same) else "FAILED")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
