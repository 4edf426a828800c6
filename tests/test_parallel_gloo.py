"""CPU tests of the multi-process host logic (world_size 2, gloo): gradient all-reduce helper,
session partitioning and the invariance of the negative sampler's counter stream under sharding.
The CUDA kernels themselves are single-GPU; these tests cover the plumbing around them."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from etpgt_b200 import parallel

        assert parallel.world() == (rank, world)
        torch.manual_seed(0)
        small = [torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(5))]
        large = torch.nn.Parameter(torch.zeros(1100, 1024))          # >= 4 MB: its own all-reduce
        skipped = torch.nn.Parameter(torch.zeros(3))                 # no gradient: must be ignored
        for p in small + [large]:
            p.grad = torch.full_like(p, float(rank + 1))
        parallel.allreduce_gradients(small + [large, skipped])
        ok = all(bool((p.grad == 3.0).all()) for p in small + [large]) and skipped.grad is None
        # sessions of this rank + global-index Philox stream (oracle restatement of the device sampler)
        from oracle import graph_ref

        rng = np.random.default_rng(1)
        sessions = [rng.integers(1, 50, size=rng.integers(3, 9)) for _ in range(40)]
        cost = np.array([len(s) for s in sessions])
        cuts = parallel.partition_sessions(cost, world)
        mine = range(cuts[rank], cuts[rank + 1])
        neg = {s: graph_ref.sample_negatives(11, 0, s, sessions[s], 60, 5).tolist() for s in mine}
        gathered = [None] * world
        dist.all_gather_object(gathered, neg)
        if rank == 0:
            merged = {k: v for part in gathered for k, v in part.items()}
            single = {s: graph_ref.sample_negatives(11, 0, s, sessions[s], 60, 5).tolist() for s in range(40)}
            ok = ok and merged == single
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_two_gloo():
    world = 2
    port = _free_port()
    manager = mp.get_context("spawn").Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_loader_shares_cover_every_global_batch_exactly_once():
    """Host logic of the sharding loader (DeviceSessionLoader._share): contiguous shares balanced by session
    length, at least two sessions per rank, tiny batches replicated on every rank."""
    from etpgt_b200.train.dataloader import DeviceSessionLoader

    rng = np.random.default_rng(0)
    lengths = rng.integers(2, 50, size=1000).astype(np.float64)
    for world in (2, 3, 8):
        for size in (5, 16, 17, 64, 333):
            ids = rng.permutation(1000)[:size]
            shares, counts_seen = [], None
            for rank in range(world):
                loader = DeviceSessionLoader.__new__(DeviceSessionLoader)
                loader._lengths, loader.rank, loader.world_size = lengths, rank, world
                mine, counts, replicated = loader._share(ids)
                assert counts_seen in (None, counts)
                counts_seen = counts
                assert len(mine) == counts[rank]
                shares.append((mine, replicated))
            if size < 2 * world:
                assert all(rep and np.array_equal(m, ids) for m, rep in shares)
            else:
                assert not any(rep for _, rep in shares)
                assert np.array_equal(np.concatenate([m for m, _ in shares]), ids)
                assert min(len(m) for m, _ in shares) >= 2
                cost = [lengths[m].sum() for m, _ in shares]
                assert max(cost) - min(cost) <= 2 * lengths.max() + 1e-9 or size < 4 * world
