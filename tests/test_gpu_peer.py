"""Peer-memory data parallelism: the communicator kernels (barrier, small fp64 all-reduce, fp32
sum), the fused reduce-scatter + AdamW + all-gather of the item table, and the whole data-parallel training step.

The first group runs SEVERAL RANKS INSIDE ONE PROCESS on one GPU (`PeerComm.local_group`: the regions are plain
device pointers instead of IPC mappings, every rank has its own stream) — the kernels, flags and orderings are the
ones used across GPUs, so the protocol is covered wherever one B200 is available.  The second group spawns two
processes on two GPUs (CUDA IPC + NCCL for the plumbing) and is skipped on a single-GPU box.

Oracles: sums in rank order formed by torch (bit-exact: the kernels promise exactly that order); the single-GPU
optimizer kernel etpgt_adam_step on the summed gradient (bit-exact, same arithmetic; that kernel itself is checked
against torch.optim and the fp64 oracle in tests/test_gpu_optim.py); the single-process whole-batch training step
(<= 1e-5 of the largest gradient: the only difference is the summation order across the shards)."""

import os
import socket
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _group(world, extra_bytes=0):
    from etpgt_b200 import _lib
    from etpgt_b200.parallel import PeerComm

    comms = PeerComm.local_group(world, int(_lib.size("etpgt_comm_control_bytes")) + extra_bytes)
    for c in comms:
        c.set_timeout(5.0)
    return comms, [torch.cuda.Stream() for _ in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_small_allreduce_sums_in_rank_order_and_reuses_its_slots(world):
    comms, streams = _group(world)
    g = torch.Generator(device="cuda").manual_seed(world)
    rounds, count = 21, 513          # 21 > 2 * the slot ring: every slot is reused at least twice
    inputs = [[torch.randn(count, dtype=torch.float64, device="cuda", generator=g) for _ in range(world)]
              for _ in range(rounds)]
    outs = [[torch.empty(count, dtype=torch.float64, device="cuda") for _ in range(world)] for _ in range(rounds)]
    torch.cuda.synchronize()
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            for i in range(rounds):
                if i % 2:
                    comms[r].allreduce_f64(inputs[i][r], outs[i][r])
                else:                 # in place
                    outs[i][r].copy_(inputs[i][r])
                    comms[r].allreduce_f64(outs[i][r])
            comms[r].barrier()
    torch.cuda.synchronize()
    for i in range(rounds):
        want = torch.zeros(count, dtype=torch.float64, device="cuda")
        for r in range(world):
            want = want + inputs[i][r]
        for r in range(world):
            assert torch.equal(outs[i][r], want), (i, r)
    assert [c.status() for c in comms] == [0] * world


def test_a_missing_peer_times_out_instead_of_hanging():
    comms, streams = _group(2)
    comms[0].set_timeout(0.2)
    x = torch.ones(8, dtype=torch.float64, device="cuda")
    with torch.cuda.stream(streams[0]):
        comms[0].allreduce_f64(x.clone())      # rank 1 never shows up
    torch.cuda.synchronize()
    assert comms[0].status() == 2
    with pytest.raises(RuntimeError, match="timed out"):
        comms[0].check()


def test_sum_f32_over_peer_regions():
    from etpgt_b200 import _lib

    world, numel = 4, 1_000_003
    comms, streams = _group(world, extra_bytes=4 * numel + 256)
    off = int(_lib.size("etpgt_comm_control_bytes"))
    g = torch.Generator(device="cuda").manual_seed(3)
    parts = [comms[r].tensor(off, (numel,)) for r in range(world)]
    for t in parts:
        t.copy_(torch.randn(numel, device="cuda", generator=g))
    outs = [torch.empty(numel, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            comms[r].barrier()
            comms[r].sum_f32(off, numel, outs[r])
            comms[r].barrier()
    torch.cuda.synchronize()
    want = ((parts[0] + parts[1]) + parts[2]) + parts[3]
    for r in range(world):
        assert torch.equal(outs[r], want)


@pytest.mark.parametrize("world,decoupled", [(2, True), (4, False), (8, True)])
def test_fused_table_update_equals_adam_on_the_summed_gradient(world, decoupled):
    """etpgt_dp_adam_table on every rank == etpgt_adam_step on sum_r grad_r, bit for bit, in every rank's copy of
    the table; moments are current on the owned rows only."""
    from etpgt_b200 import _lib, parallel
    from etpgt_b200.optim import _AdamTensor

    rows, dim = 1003, 64             # not a multiple of the world size: the last shard is short
    nbytes = 4 * rows * dim
    pad = (nbytes + 255) // 256 * 256
    comms, streams = _group(world, extra_bytes=2 * pad)
    off_g = int(_lib.size("etpgt_comm_control_bytes"))
    off_p = off_g + pad
    g = torch.Generator(device="cuda").manual_seed(11)
    p0 = torch.randn(rows, dim, device="cuda", generator=g)
    m0 = torch.randn(rows, dim, device="cuda", generator=g) * 0.1
    v0 = torch.rand(rows, dim, device="cuda", generator=g) * 0.01
    grads = [torch.randn(rows, dim, device="cuda", generator=g) for _ in range(world)]
    tables = [comms[r].tensor(off_p, (rows, dim)) for r in range(world)]
    sinks = [comms[r].tensor(off_g, (rows, dim)) for r in range(world)]
    moments = [(m0.clone(), v0.clone()) for _ in range(world)]
    for r in range(world):
        tables[r].copy_(p0)
        sinks[r].copy_(grads[r])
    hyper = (1e-3, 0.9, 0.999, 1e-8, 1e-2, int(decoupled), 7)
    torch.cuda.synchronize()
    for r in range(world):
        lo, hi = parallel.item_shard(rows, r, world)
        with torch.cuda.stream(streams[r]):
            comms[r].barrier()
            _lib.call("etpgt_dp_adam_table", comms[r].handle, off_p, off_g, _lib.ptr(moments[r][0]),
                      _lib.ptr(moments[r][1]), rows, dim, lo, hi, *[float(h) for h in hyper[:5]], hyper[5], hyper[6],
                      _lib.stream())
            comms[r].barrier()
    torch.cuda.synchronize()
    total = grads[0].clone()
    for r in range(1, world):
        total = total + grads[r]
    want_p, want_m, want_v = p0.clone(), m0.clone(), v0.clone()
    arr = (_AdamTensor * 1)(_AdamTensor(want_p.data_ptr(), total.data_ptr(), want_m.data_ptr(), want_v.data_ptr(),
                                        want_p.numel()))
    _lib.call("etpgt_adam_step", arr, 1, *[float(h) for h in hyper[:5]], hyper[5], hyper[6], 0, _lib.stream())
    torch.cuda.synchronize()
    for r in range(world):
        assert torch.equal(tables[r], want_p), f"table copy of rank {r}"
        lo, hi = parallel.item_shard(rows, r, world)
        assert torch.equal(moments[r][0][lo:hi], want_m[lo:hi]) and torch.equal(moments[r][1][lo:hi], want_v[lo:hi])
        assert torch.equal(moments[r][0][:lo], m0[:lo]) and torch.equal(moments[r][0][hi:], m0[hi:])
        assert torch.equal(sinks[r], grads[r])           # the kernel does not clear gradients (the caller does)
    assert [c.status() for c in comms] == [0] * world


def _replicas(world, dim=64, dropout=0.0, num_items=500):
    """`world` model replicas in this process (identical parameters), each with its own peer region / stream."""
    from etpgt_b200 import optim, parallel
    from etpgt_b200.model import create_graph_transformer_optimized

    def make():
        torch.manual_seed(0)
        model = create_graph_transformer_optimized(num_items, dim, dim, dropout=dropout, laplacian_k=8).cuda()
        model.laplacian_pe._cached_pe = torch.randn(num_items, 8, generator=torch.Generator().manual_seed(7)).abs().cuda()
        return model.train()

    models = [make() for _ in range(world)]
    table = models[0].item_embedding.weight
    dense = sum(p.numel() for p in models[0].parameters()) - table.numel()
    from etpgt_b200 import _lib
    need = int(_lib.size("etpgt_comm_control_bytes")) + 4 * dense + 8 * table.numel() + 4096
    comms = parallel.PeerComm.local_group(world, need)
    peers = []
    for model, comm in zip(models, comms):
        comm.set_timeout(5.0)
        peers.append(parallel.PeerDataParallel(model, comm=comm))
    opts = [optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5) for m in models]
    return models, peers, opts, make


@pytest.mark.parametrize("world", [2, 3])
def test_data_parallel_step_over_peer_memory_matches_the_whole_batch(world):
    """Ranks in one process, contiguous session shares of one global batch: the exchanged gradients equal the
    single-process gradient of the whole batch (summation order is the only difference), the replicas stay
    bit-identical through optimizer steps, and a second run reproduces the first bit for bit."""
    from etpgt_b200 import data, ops, optim, parallel, synth
    from etpgt_b200.train.step import FusedTrainStep

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    global_batch, steps = 384, 3

    def shares(step):
        ids = np.arange(step * global_batch, (step + 1) * global_batch) % d.num_sessions
        cost = (d.sess_ptr[ids + 1] - d.sess_ptr[ids]).astype(np.float64)
        cuts = parallel.partition_sessions(cost, world)
        return ids, [ids[cuts[r]:cuts[r + 1]] for r in range(world)]

    def batch_of(ids, step):
        batch = data.build_batch(graph, store, ids)
        batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=3, step=step)
        ops.prepare_batch(batch, d.num_items)
        return batch

    def run():
        models, peers, opts, make = _replicas(world)
        fused = [FusedTrainStep(m, "bpr") for m in models]
        streams = [torch.cuda.Stream() for _ in range(world)]
        first = None
        for step in range(steps):
            _, parts = shares(step)
            batches = [batch_of(ids, step) for ids in parts]
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    opts[r].zero_grad()
                    fused[r](batches[r], total_sessions=global_batch)
            if step == 0:
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        fused[r].allreduce_gradients()
                torch.cuda.synchronize()
                first = [{n: p.grad.detach().clone() for n, p in models[r].named_parameters()
                          if n != "item_embedding.weight"} for r in range(world)]
                for r in range(world):
                    with torch.cuda.stream(streams[r]):
                        table_grad = peers[r].reduced_table_gradient()
                    first[r]["item_embedding.weight"] = table_grad
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    opts[r].step()
            torch.cuda.synchronize()
            for peer in peers:
                peer.comm.check()
        return models, first, make

    models, first, make = run()
    # (a) replicas are bit-identical after the steps (parameters and BatchNorm running statistics)
    ref_state = models[0].state_dict()
    for r in range(1, world):
        for k, v in models[r].state_dict().items():
            assert torch.equal(v, ref_state[k]), f"rank {r}: {k}"
    # (b) the exchanged step-0 gradient == the single-process gradient of the whole batch
    single = make()
    single.bn_process_group = False
    ids, _ = shares(0)
    batch = batch_of(ids, 0)
    loss = ops.sampled_loss(single(batch), single.item_embedding, batch.target_item, batch.negative_items, "bpr",
                            total_sessions=global_batch)[0]
    loss.backward()
    want = {n: p.grad.detach() for n, p in single.named_parameters()}
    scale = max(g.double().abs().max().item() for g in want.values())
    for r in range(world):
        for name, g in want.items():
            err = (first[r][name].double() - g.double()).abs().max().item() / scale
            assert err < 1e-5, f"rank {r} {name}: {err:.3e}"
    for name in want:                      # every rank holds the SAME reduced gradient, bit for bit
        for r in range(1, world):
            assert torch.equal(first[r][name], first[0][name]), name
    # (c) deterministic: a second run gives the same bits
    again, _, _ = run()
    for k, v in again[0].state_dict().items():
        assert torch.equal(v, ref_state[k]), k


def test_peer_step_with_dropout_keeps_replicas_identical_and_finite():
    from etpgt_b200 import data, ops, synth
    from etpgt_b200.train.step import FusedTrainStep

    world = 2
    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    models, peers, opts, _ = _replicas(world, dropout=0.1)
    fused = [FusedTrainStep(m, "dual") for m in models]
    streams = [torch.cuda.Stream() for _ in range(world)]
    for step in range(2):
        batches = []
        for r in range(world):
            ids = np.arange(100 * r, 100 * r + 100) + 200 * step
            batch = data.build_batch(graph, store, ids)
            batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=step)
            ops.prepare_batch(batch, d.num_items)
            batches.append(batch)
        torch.cuda.synchronize()
        losses = []
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                opts[r].zero_grad()
                losses.append(fused[r](batches[r], total_sessions=200))
                opts[r].step()
        torch.cuda.synchronize()
        assert all(torch.isfinite(l).all() for l in losses)
    for k, v in models[1].state_dict().items():
        assert torch.equal(v, models[0].state_dict()[k]), k
    for peer in peers:
        peer.comm.check()


# ------------------------------------------------------------------------------ two processes, two GPUs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _two_gpu_worker(rank, world, port, out):
    import torch.distributed as dist

    import faulthandler

    # a worker that hangs (a collective only one rank entered, ...) dumps its Python stacks and exits instead of
    # hanging the suite
    faulthandler.dump_traceback_later(300, exit=True)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from etpgt_b200 import data, ops, optim, parallel, synth
        from etpgt_b200.model import create_graph_transformer_optimized
        from etpgt_b200.train.step import FusedTrainStep

        d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
        graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
        store = data.SessionStore(d.sess_ptr, d.sess_items)
        global_batch = 384

        def make():
            torch.manual_seed(0)
            m = create_graph_transformer_optimized(d.num_items, 64, 64, dropout=0.0, laplacian_k=8).cuda()
            m.laplacian_pe._cached_pe = torch.randn(d.num_items, 8, generator=torch.Generator().manual_seed(7)).abs().cuda()
            return m.train()

        def batch_of(ids, step):
            batch = data.build_batch(graph, store, ids)
            batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=3, step=step)
            ops.prepare_batch(batch, d.num_items)
            return batch

        model = make()
        peer = parallel.enable_data_parallel(model, exchange="peer")
        peer.comm.set_timeout(20.0)
        opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
        fused = FusedTrainStep(model, "bpr")
        ok, notes = True, []
        for step in range(3):
            ids = np.arange(step * global_batch, (step + 1) * global_batch) % d.num_sessions
            cost = (d.sess_ptr[ids + 1] - d.sess_ptr[ids]).astype(np.float64)
            cuts = parallel.partition_sessions(cost, world)
            opt.zero_grad()
            fused(batch_of(ids[cuts[rank]:cuts[rank + 1]], step), total_sessions=global_batch)
            if step == 0:
                fused.allreduce_gradients()
                got = {n: p.grad.detach().clone() for n, p in model.named_parameters() if n != "item_embedding.weight"}
                got["item_embedding.weight"] = peer.reduced_table_gradient()
                single = make()
                single.bn_process_group = False
                whole = batch_of(ids, 0)
                ops.sampled_loss(single(whole), single.item_embedding, whole.target_item, whole.negative_items, "bpr",
                                 total_sessions=global_batch)[0].backward()
                want = {n: p.grad.detach() for n, p in single.named_parameters()}
                scale = max(g.double().abs().max().item() for g in want.values())
                worst = max((got[n].double() - g.double()).abs().max().item() / scale for n, g in want.items())
                notes.append(f"step-0 gradient vs whole batch: {worst:.2e}")
                ok = ok and worst < 1e-5
            opt.step()
        peer.comm.check()
        # replicas bit-identical across the two GPUs
        digest = torch.stack([p.detach().double().sum() for p in model.parameters()] +
                             [b.double().sum() for b in model.buffers()])
        both = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        same = bool(torch.equal(both[0], both[1]))
        # item-sharded evaluation == single-GPU scoring
        model.eval()
        with torch.no_grad():
            sess = model(batch_of(np.arange(600 + 50 * rank, 650 + 50 * rank), 9))
            # exact for either scorer: the fp32 CUDA-core one and the tcgen05 bf16 one
            top_equal = True
            for precision in ("fp32", "bf16"):
                top = parallel.sharded_predict(model, sess, k=20, precision=precision)
                single_top = ops.score_topk(sess, model.get_item_embeddings(), 20, precision=precision)[1]
                top_equal = top_equal and bool(torch.equal(top, single_top))
        # the Trainer as one rank of a data-parallel job (process_group=True) vs. a single-process Trainer on the whole
        # batches: same epoch loss (global-batch mean), same Recall / NDCG (item-sharded scoring, counters summed)
        import tempfile

        from etpgt_b200.train.trainer import Trainer

        def loaders(shard):
            def make_list(first, count, size, base_step):
                batches = []
                for i in range(count):
                    ids = np.arange(first + i * size, first + (i + 1) * size)
                    mine, counts = ids, (size,)
                    if shard:
                        cost = (d.sess_ptr[ids + 1] - d.sess_ptr[ids]).astype(np.float64)
                        cuts = parallel.partition_sessions(cost, world)
                        mine, counts = ids[cuts[rank]:cuts[rank + 1]], tuple(int(c) for c in np.diff(cuts))
                    b = batch_of(mine, base_step + i)
                    b.negative_items = b.negative_items.reshape(-1)
                    b.total_sessions, b.rank_sessions, b.replicated = size, counts, False
                    batches.append(b)
                return batches
            return make_list(0, 3, 128, 100), make_list(600, 2, 64, 200)

        def run_trainer(dp):
            m = make()
            m.dropout_layer.p = 0.0
            if not dp:
                m.bn_process_group = False     # a single-process replica inside an initialised process group
            o = optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)      # created BEFORE the trainer re-homes the table
            train_l, val_l = loaders(dp)
            with tempfile.TemporaryDirectory() as tmp:
                t = Trainer(m, train_l, val_l, o, output_dir=Path(tmp) / f"r{rank}", max_epochs=2, patience=5,
                            k_values=[10, 20], process_group=True if dp else None)
                hist = t.train()
                wrote = (Path(tmp) / f"r{rank}" / "checkpoint_latest.pt").exists()
            return hist, wrote

        hist_dp, wrote = run_trainer(True)
        trainer_ok = wrote == (rank == 0)
        if rank == 0:
            hist_single, _ = run_trainer(False)
            for a, b in zip(hist_dp["train_loss"], hist_single["train_loss"]):
                trainer_ok = trainer_ok and abs(a - b) <= 1e-4 * abs(b)
            for ma, mb in zip(hist_dp["val_metrics"], hist_single["val_metrics"]):
                for key in mb:
                    trainer_ok = trainer_ok and abs(ma[key] - mb[key]) <= 0.02
            notes.append(f"trainer dp {hist_dp} single {hist_single}")
        out[rank] = (ok and trainer_ok, same, top_equal, notes)
    finally:
        faulthandler.cancel_dump_traceback_later()
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpus_peer_exchange_matches_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    manager = mp.get_context("spawn").Manager()
    out = manager.dict()
    mp.spawn(_two_gpu_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in (0, 1):
        ok, same, top_equal, notes = out[rank]
        assert ok and same and top_equal, (rank, ok, same, top_equal, notes)
