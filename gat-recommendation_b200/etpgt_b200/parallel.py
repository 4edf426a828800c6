"""Session-batch data parallelism over one NVLink/NVSwitch box: one process per GPU,
`torch.distributed` (NCCL) for the plumbing (the reference is single-GPU).

  * sessions are independent graph components -> contiguous session ranges per rank, balanced by
    a (nodes + edges)-like cost prefix sum, no data-path collective for the graph kernels;
  * the only couplings are (a) BatchNorm statistics — all-reduced inside ops.BatchNormRows
    (2*dim doubles per layer and direction) so whole-batch semantics are kept, (b) the loss mean
    over the GLOBAL batch, (c) the parameter gradients — two flat all-reduces (dense parameters,
    item table);
  * evaluation shards the item table by contiguous id ranges: every rank scores all sessions
    against its shard with the fused top-k kernel, candidates are all-gathered and merged exactly
    (score desc, id asc), so the result is identical for any GPU count.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_sessions(cost: np.ndarray, parts: int) -> np.ndarray:
    """Boundaries [parts+1] of contiguous session ranges with near-equal total cost
    (cost[s] ~ nodes + edges of session s)."""
    total = np.concatenate([[0], np.cumsum(cost, dtype=np.float64)])
    targets = total[-1] * np.arange(1, parts) / parts
    cuts = np.searchsorted(total, targets, side="left")
    return np.concatenate([[0], cuts, [len(cost)]]).astype(np.int64)


def item_shard(num_items: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous id range [lo, hi) of the item-table shard of `rank`."""
    per = (num_items + world_size - 1) // world_size
    lo = min(rank * per, num_items)
    return lo, min(lo + per, num_items)


def enable_global_batch_norm(model, group=None) -> None:
    """BatchNorm statistics (forward and backward) are summed over every rank's node rows."""
    model.bn_process_group = group if group is not None else dist.group.WORLD


def allreduce_gradients(params, group=None) -> None:
    """Sums gradients across ranks: one flat all-reduce for the dense parameters and one for each
    large (>= 4 MB) gradient such as the item table, in place.  The loss is already divided by the
    global batch, so SUM gives the global-batch gradient."""
    rank, size = world()
    if size == 1:
        return
    small, large = [], []
    for p in params:
        if p.grad is None:
            continue
        (large if p.grad.numel() * p.grad.element_size() >= (4 << 20) else small).append(p)
    if small:
        flat = torch.cat([p.grad.reshape(-1) for p in small])
        dist.all_reduce(flat, group=group)
        offset = 0
        for p in small:   # rebind .grad to its slice of the reduced buffer: no copy back
            n = p.grad.numel()
            p.grad = flat[offset:offset + n].view_as(p.grad)
            offset += n
    for p in large:
        dist.all_reduce(p.grad, group=group)


@torch.no_grad()
def sharded_predict(model, session_embeddings: torch.Tensor, k: int = 20, group=None) -> torch.Tensor:
    """Item-sharded full-catalogue top-k: all-gather the session vectors, score the local id range,
    all-gather (value, id) candidates, exact merge.  Returns the top-k ids of THIS rank's sessions."""
    from . import ops

    rank, size = world()
    table = model.get_item_embeddings()
    if size == 1:
        return ops.score_topk(session_embeddings, table, k)[1]
    counts = [torch.zeros(1, dtype=torch.int64, device=session_embeddings.device) for _ in range(size)]
    dist.all_gather(counts, torch.tensor([session_embeddings.size(0)], device=session_embeddings.device), group=group)
    counts = [int(c.item()) for c in counts]
    width = session_embeddings.size(1)
    padded = torch.zeros(max(counts), width, dtype=torch.float32, device=session_embeddings.device)
    padded[: session_embeddings.size(0)] = session_embeddings
    gathered = [torch.empty_like(padded) for _ in range(size)]
    dist.all_gather(gathered, padded, group=group)
    everyone = torch.cat([g[:c] for g, c in zip(gathered, counts)])
    lo, hi = item_shard(table.size(0), rank, size)
    kk = min(k, hi - lo)
    val, idx = ops.score_topk(everyone, table[lo:hi], kk, id_base=lo)
    if kk < k:  # tiny shard: pad with sentinels that lose every comparison
        pad_v = torch.full((val.size(0), k - kk), float("-inf"), device=val.device)
        pad_i = torch.full((val.size(0), k - kk), torch.iinfo(torch.int64).max, device=val.device)
        val, idx = torch.cat([val, pad_v], 1), torch.cat([idx, pad_i], 1)
    vals = [torch.empty_like(val) for _ in range(size)]
    idxs = [torch.empty_like(idx) for _ in range(size)]
    dist.all_gather(vals, val, group=group)
    dist.all_gather(idxs, idx, group=group)
    start = sum(counts[:rank])
    mine = slice(start, start + counts[rank])
    cand_v = torch.cat([v[mine] for v in vals], dim=1).contiguous()
    cand_i = torch.cat([i[mine] for i in idxs], dim=1).contiguous()
    return ops.topk_merge(cand_v, cand_i, k)[1]
