"""TEST INFRASTRUCTURE ONLY — numpy / pure-Python oracle for the integer side of the hot path
(SURVEY.md §8a rows a1-a3 and the index structures the kernels consume).  Every output here
is compared BIT-EXACTLY with the CUDA path.

  * `csr_from_coo`            — the destination-sorted CSR + source-sorted CSC the fused
                                convolutions walk (no reference counterpart: PyG scatters over
                                COO; the contract is "stable by original edge index").
  * `session_subgraph` / `collate_sessions`
                              — etpgt/train/dataloader.py:126-202 (train_baseline rule) and
                                scripts/pipeline/run_full_pipeline.py:120-149 (pipeline rule).
  * `philox4x32_10`, `sample_negatives`
                              — etpgt/train/dataloader.py:107-124 acceptance rule driven by a
                                counter-based Philox stream (the reference's own mt19937 stream
                                is worker-order dependent).
  * `merge_topk`              — exact (score desc, id asc) merge of per-shard candidates.
  * `co_event_graph`          — scripts/data/04_build_graph.py:25-127 (`build_co_event_graph`), the
                                step before the path: window-5 pair counts, last
                                timestamp, edges ordered by count descending with ties in dict-insertion
                                (first emission) order.  Pinned against the reference function itself
                                through tests/golden/co_event_graph.npz.
"""

from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- CSR / CSC


def csr_from_coo(src: np.ndarray, dst: np.ndarray, num_nodes: int):
    """Returns dict(rowptr, col, eperm, colptr, row, cpos).

    CSR: edges stably sorted by destination; col[p] = source, eperm[p] = original edge index.
    CSC: the CSR-ordered edges stably sorted by source; row[p] = destination, cpos[p] = the
    edge's position in CSR order."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    eperm = np.argsort(dst, kind="stable")
    col = src[eperm]
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    cpos = np.argsort(col, kind="stable")
    row = dst[eperm][cpos]
    colptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(colptr, src + 1, 1)
    colptr = np.cumsum(colptr)
    as32 = lambda a: a.astype(np.int32)  # noqa: E731
    return dict(rowptr=as32(rowptr), col=as32(col), eperm=as32(eperm),
                colptr=as32(colptr), row=as32(row), cpos=as32(cpos))


# ----------------------------------------------------------------------------- sessions


def build_item_graph(item_i: np.ndarray, item_j: np.ndarray, num_items: int):
    """Lookup structure over the global co-occurrence edge list in its stored (CSV) order:
    rows keyed by item_i, each row sorted by item_j, with the CSV row index as payload.
    (The reference scans the whole frame per sample instead, dataloader.py:138-140.)"""
    item_i = np.asarray(item_i, dtype=np.int64)
    item_j = np.asarray(item_j, dtype=np.int64)
    order = np.lexsort((item_j, item_i))
    gptr = np.zeros(num_items + 1, dtype=np.int64)
    np.add.at(gptr, item_i + 1, 1)
    return dict(gptr=np.cumsum(gptr), gcol=item_j[order], gidx=order.astype(np.int64))


def session_context(items: np.ndarray, max_len: int = 50):
    """dataloader.py:84-92: keep the last `max_len` events, target = last, context = rest."""
    items = np.asarray(items, dtype=np.int64)
    if len(items) > max_len:
        items = items[-max_len:]
    return items[:-1], int(items[-1]), items


def session_subgraph(context: np.ndarray, item_i: np.ndarray, item_j: np.ndarray,
                     symmetrize: bool = False, self_loop_if_empty: bool = False):
    """Nodes = sorted unique context items (dataloader.py:173 `unique()`); an edge of the global
    list is kept iff both ends are context items, in stored order and direction
    (dataloader.py:138-152).  `symmetrize` appends the reversed copies after the forward ones
    and `self_loop_if_empty` adds one loop per node when nothing was kept
    (run_full_pipeline.py:143-149)."""
    nodes = np.unique(np.asarray(context, dtype=np.int64))
    keep = np.isin(item_i, nodes) & np.isin(item_j, nodes)
    src = np.searchsorted(nodes, item_i[keep])
    dst = np.searchsorted(nodes, item_j[keep])
    if symmetrize:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    if self_loop_if_empty and len(src) == 0:
        src = dst = np.arange(len(nodes), dtype=np.int64)
    return nodes, src.astype(np.int64), dst.astype(np.int64)


def collate_sessions(sessions: list[np.ndarray], item_i, item_j, max_len: int = 50,
                     symmetrize: bool = False, self_loop_if_empty: bool = False):
    """dataloader.py:157-202 + PyG `Batch.from_data_list`: concatenated node ids, edges shifted by
    the cumulative node count, graph id per node, one target per session."""
    xs, srcs, dsts, batch, targets = [], [], [], [], []
    node_ptr, edge_ptr = [0], [0]
    for s, items in enumerate(sessions):
        ctx, target, _ = session_context(items, max_len)
        nodes, src, dst = session_subgraph(ctx, item_i, item_j, symmetrize, self_loop_if_empty)
        xs.append(nodes)
        srcs.append(src + node_ptr[-1])
        dsts.append(dst + node_ptr[-1])
        batch.append(np.full(len(nodes), s, dtype=np.int64))
        targets.append(target)
        node_ptr.append(node_ptr[-1] + len(nodes))
        edge_ptr.append(edge_ptr[-1] + len(src))
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)  # noqa: E731
    return dict(x=cat(xs), edge_src=cat(srcs), edge_dst=cat(dsts), batch=cat(batch),
                target=np.asarray(targets, dtype=np.int64),
                node_ptr=np.asarray(node_ptr, dtype=np.int64), edge_ptr=np.asarray(edge_ptr, dtype=np.int64))


# ----------------------------------------------------------------------------- Philox sampler

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32 with 10 rounds (Salmon et al., SC'11): counter (4 x u32), key (2 x u32)."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def sample_negatives(seed: int, step: int, session_index: int, session_items, num_items: int, num_neg: int):
    """`num_neg` uniform ids in [1, num_items) that are not session items (target included);
    duplicates among the negatives are allowed (dataloader.py:116-124).

    Stream: key = (seed lo, seed hi); counter = (GLOBAL session index, slot, attempt // 4, step);
    attempt a uses word a % 4; candidate = 1 + ((word * (num_items - 1)) >> 32)."""
    members = {int(v) for v in session_items}
    key = (seed & _MASK, (seed >> 32) & _MASK)
    out = []
    for slot in range(num_neg):
        attempt = 0
        while True:
            words = philox4x32_10((session_index, slot, attempt // 4, step), key)
            cand = 1 + ((words[attempt % 4] * (num_items - 1)) >> 32)
            attempt += 1
            if cand not in members:
                out.append(cand)
                break
    return np.asarray(out, dtype=np.int64)


# ----------------------------------------------------------------------------- top-k merge


def merge_topk(values: np.ndarray, ids: np.ndarray, k: int):
    """values/ids [B, M] candidate lists (any order) -> the k best by (score desc, id asc)."""
    out_v = np.empty((values.shape[0], k), dtype=values.dtype)
    out_i = np.empty((values.shape[0], k), dtype=ids.dtype)
    for b in range(values.shape[0]):
        order = np.lexsort((ids[b], -values[b].astype(np.float64)))[:k]
        out_v[b], out_i[b] = values[b][order], ids[b][order]
    return out_v, out_i


def co_event_graph(sess_ptr, sess_items, timestamps=None, window: int = 5):
    """scripts/data/04_build_graph.py:41-101 on sessions already grouped and time-sorted.
    Returns (item_i, item_j, count, last_ts) int64 arrays, count-descending, ties by first emission
    (a STABLE sort of the reference's dict order; pandas' default quicksort leaves ties unspecified)."""
    edges: dict[tuple[int, int], list[int]] = {}
    sess_ptr = np.asarray(sess_ptr)
    sess_items = np.asarray(sess_items)
    for s in range(len(sess_ptr) - 1):
        lo, hi = int(sess_ptr[s]), int(sess_ptr[s + 1])
        for i in range(lo, hi):
            for j in range(i + 1, min(i + window + 1, hi)):
                a, b = int(sess_items[i]), int(sess_items[j])
                if a > b:                                   # 04_build_graph.py:64-71
                    a, b = b, a
                    ts = int(timestamps[j]) if timestamps is not None else 0
                else:
                    ts = int(timestamps[i]) if timestamps is not None else 0
                rec = edges.setdefault((a, b), [0, 0])
                rec[0] += 1
                rec[1] = max(rec[1], ts)
    keys = list(edges)
    order = sorted(range(len(keys)), key=lambda r: -edges[keys[r]][0])   # stable
    item_i = np.asarray([keys[r][0] for r in order], dtype=np.int64)
    item_j = np.asarray([keys[r][1] for r in order], dtype=np.int64)
    count = np.asarray([edges[keys[r]][0] for r in order], dtype=np.int64)
    last_ts = np.asarray([edges[keys[r]][1] for r in order], dtype=np.int64)
    return item_i, item_j, count, last_ts
