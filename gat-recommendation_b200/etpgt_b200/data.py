"""GPU-resident data path: the per-batch work of the reference's SessionDataset / collate_fn
(etpgt/train/dataloader.py:12-202) done by device kernels over data that lives in HBM.

    graph    = ItemGraph(item_i, item_j, num_items)            # graph_edges.csv columns, once
    sessions = SessionStore(sess_ptr, sess_items)               # train.csv grouped by session, once
    batch    = build_batch(graph, sessions, session_ids)        # x / edge_index / batch / target_item
    batch.negative_items = sample_negatives(sessions, session_ids, graph.num_items, 5, seed, step)

`batch` has the attributes of the PyG Batch the reference hands to `model(batch)`.
"""

from __future__ import annotations

import torch

from ._lib import call, ptr, size, stream, workspace


def _dev_i64(t, device) -> torch.Tensor:
    t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.int64).contiguous()


class ItemGraph:
    """Lookup structure over the stored co-occurrence edge list (rows keyed by item_i, sorted by
    item_j, payload = stored row index), built on the device."""

    def __init__(self, item_i, item_j, num_items: int, device="cuda"):
        item_i, item_j = _dev_i64(item_i, device), _dev_i64(item_j, device)
        e = item_i.numel()
        self.num_items, self.num_edges = int(num_items), e
        i32 = dict(dtype=torch.int32, device=item_i.device)
        self.gptr = torch.empty(num_items + 1, **i32)
        self.gcol = torch.empty(max(e, 1), **i32)
        self.gidx = torch.empty(max(e, 1), **i32)
        ws = workspace(size("etpgt_item_graph_workspace_bytes", e), item_i.device)
        call("etpgt_item_graph_build", ptr(item_i), ptr(item_j), e, num_items, ptr(self.gptr), ptr(self.gcol),
             ptr(self.gidx), ptr(ws), ws.numel(), stream())


class SessionStore:
    """All sessions, resident on the device: sess_items[sess_ptr[s]:sess_ptr[s+1]] in time order."""

    def __init__(self, sess_ptr, sess_items, device="cuda"):
        self.ptr = _dev_i64(sess_ptr, device)
        self.items = _dev_i64(sess_items, device)
        self.num_sessions = self.ptr.numel() - 1


class SessionBatch:
    """Duck-typed PyG Batch: x, edge_index, batch, ptr, target_item, negative_items, num_graphs."""

    def __init__(self, x, edge_index, batch, node_ptr, target_item, num_graphs):
        self.x, self.edge_index, self.batch, self.ptr = x, edge_index, batch, node_ptr
        self.target_item, self.num_graphs = target_item, num_graphs
        self.negative_items = None

    @property
    def num_nodes(self) -> int:
        return int(self.x.numel())

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.size(1))

    def to(self, device):
        for name in ("x", "edge_index", "batch", "ptr", "target_item", "negative_items"):
            value = getattr(self, name)
            if torch.is_tensor(value):
                setattr(self, name, value.to(device))
        return self


def build_batch(graph: ItemGraph, sessions: SessionStore, session_ids=None, max_len: int = 50,
                symmetrize: bool = False, self_loop_if_empty: bool = False) -> SessionBatch:
    """`symmetrize=False, self_loop_if_empty=False` is the train_baseline rule (dataloader.py:126-154);
    both True is the run_full_pipeline rule (run_full_pipeline.py:143-149)."""
    dev = sessions.ptr.device
    ids = None if session_ids is None else _dev_i64(session_ids, dev)
    b = sessions.num_sessions if ids is None else ids.numel()
    node_ptr = torch.empty(b + 1, dtype=torch.int32, device=dev)
    edge_ptr = torch.empty(b + 1, dtype=torch.int32, device=dev)
    ws = workspace(size("etpgt_session_subgraphs_workspace_bytes", b, 0), dev)
    call("etpgt_session_subgraphs_count", ptr(graph.gptr), ptr(graph.gcol), ptr(sessions.ptr), ptr(sessions.items),
         ptr(ids), b, max_len, int(symmetrize), int(self_loop_if_empty), ptr(node_ptr), ptr(edge_ptr), ptr(ws),
         ws.numel(), stream())
    # the one host read of the data path: the two totals that size the batch tensors
    n, e = torch.stack([node_ptr[-1], edge_ptr[-1]]).tolist()
    i64 = dict(dtype=torch.int64, device=dev)
    x, bvec = torch.empty(n, **i64), torch.empty(n, **i64)
    edge_index = torch.empty(2, e, **i64)
    target = torch.empty(b, **i64)
    ws = workspace(size("etpgt_session_subgraphs_workspace_bytes", b, e), dev)
    call("etpgt_session_subgraphs_fill", ptr(graph.gptr), ptr(graph.gcol), ptr(graph.gidx), ptr(sessions.ptr),
         ptr(sessions.items), ptr(ids), b, max_len, int(symmetrize), int(self_loop_if_empty), ptr(node_ptr),
         ptr(edge_ptr), e, ptr(x), ptr(bvec), ptr(edge_index[0]), ptr(edge_index[1]), ptr(target), ptr(ws),
         ws.numel(), stream())
    return SessionBatch(x, edge_index, bvec, node_ptr, target, b)


def sample_negatives(sessions: SessionStore, session_ids, num_items: int, num_neg: int = 5, seed: int = 0,
                     step: int = 0, session_base: int = 0, max_len: int = 50) -> torch.Tensor:
    """[B, num_neg] negatives, never a session item, never the padding id 0 (dataloader.py:107-124);
    Philox counters are keyed by the GLOBAL session index, so any sharding gives the same ids."""
    dev = sessions.ptr.device
    ids = None if session_ids is None else _dev_i64(session_ids, dev)
    b = sessions.num_sessions if ids is None else ids.numel()
    out = torch.empty(b, num_neg, dtype=torch.int64, device=dev)
    call("etpgt_sample_negatives", int(seed) & (2 ** 64 - 1), int(step) & 0xFFFFFFFF, int(session_base),
         ptr(sessions.ptr), ptr(sessions.items), ptr(ids), b, max_len, num_items, num_neg, ptr(out), stream())
    return out


def build_co_event_graph(sess_ptr, sess_items, timestamps=None, window: int = 5, num_items: int | None = None,
                         device="cuda"):
    """The co-occurrence graph of scripts/data/04_build_graph.py:25-127 built on the device
    (`etpgt_cooc_graph_build`): sessions given as ptr / items (/ timestamps) in time order.  Returns
    device tensors (item_i, item_j, count, last_ts) ordered by count descending (ties: first emission) —
    the rows of graph_edges.csv minus the unused event_pair_hist column."""
    ptr_d, items_d = _dev_i64(sess_ptr, device), _dev_i64(sess_items, device)
    ts_d = None if timestamps is None else _dev_i64(timestamps, device)
    s, t = ptr_d.numel() - 1, items_d.numel()
    if num_items is None:
        num_items = int(items_d.max().item()) + 1 if t else 1
    cap = max(t * window, 1)
    i64 = dict(dtype=torch.int64, device=ptr_d.device)
    item_i, item_j, count = torch.empty(cap, **i64), torch.empty(cap, **i64), torch.empty(cap, **i64)
    last_ts = torch.empty(cap, **i64) if ts_d is not None else None
    num_edges = torch.zeros(1, **i64)
    ws = workspace(size("etpgt_cooc_graph_workspace_bytes", t, window), ptr_d.device)
    call("etpgt_cooc_graph_build", ptr(ptr_d), ptr(items_d), ptr(ts_d), s, t, window, num_items, cap, ptr(item_i),
         ptr(item_j), ptr(count), ptr(last_ts), ptr(num_edges), ptr(ws), ws.numel(), stream())
    e = int(num_edges.item())          # the one host read of this one-off build
    out = [item_i[:e].clone(), item_j[:e].clone(), count[:e].clone()]
    out.append(last_ts[:e].clone() if last_ts is not None else None)
    return tuple(out)
