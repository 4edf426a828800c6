"""Synthetic RetailRocket-shaped data (host side, setup only — nothing here is on the timed path).

Follows the semantics of the reference's generator and ETL (scripts/data/00_generate_synthetic_data.py:24-139
zipf popularity + revisits; 02_sessionize min length 3; 04_build_graph.py:25-127 window-5
co-occurrence, canonical item_i <= item_j, self pairs kept, rows sorted by count descending) with a
session-length law matched to docs/DATA_PIPELINE.md:127-134 (min 3, median 4, mean ~5.5, long tail
truncated at 50 by the loader).  Item ids are dense in [1, num_items); id 0 is the padding item.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SyntheticData:
    num_items: int
    item_i: np.ndarray        # [E] int64, stored (CSV) order = count descending
    item_j: np.ndarray
    sess_ptr: np.ndarray      # [S+1] int64
    sess_items: np.ndarray    # [sum len] int64, chronological inside a session

    @property
    def num_sessions(self) -> int:
        return len(self.sess_ptr) - 1

    def session(self, s: int) -> np.ndarray:
        return self.sess_items[self.sess_ptr[s]:self.sess_ptr[s + 1]]

    def stats(self) -> dict:
        lens = np.diff(self.sess_ptr)
        nodes = np.union1d(self.item_i, self.item_j)
        return {"items": int(self.num_items), "graph_nodes": int(len(nodes)), "graph_edges": int(len(self.item_i)),
                "avg_degree": float(2 * len(self.item_i) / max(len(nodes), 1)), "sessions": int(len(lens)),
                "len_mean": float(lens.mean()), "len_median": float(np.median(lens)), "len_max": int(lens.max())}


def session_lengths(rng, count: int, law: str = "retailrocket") -> np.ndarray:
    """"retailrocket": min 3, median 4, mean ~5.5, tail into the hundreds (docs/DATA_PIPELINE.md:127-134);
    "yoochoose": min 3 after the length filter, mean ~4, median 3."""
    if law == "yoochoose":
        tail = rng.pareto(2.2, size=count) * 1.7
        return np.minimum(3 + np.floor(tail).astype(np.int64), 200)
    tail = rng.pareto(1.9, size=count) * 2.65
    return np.minimum(3 + np.floor(tail).astype(np.int64), 417)


def generate(num_sessions: int = 167_705, graph_sessions: int | None = 120_436, num_items: int = 82_174,
             clusters: int = 1600, p_local: float = 0.94, a_local: float = 0.4, a_global: float = 0.5,
             a_cluster: float = 0.5, revisit: float = 0.34, window: int = 5, seed: int = 42,
             length_law: str = "retailrocket", build_graph: bool = True) -> SyntheticData:
    """Default sizes approximate RR-synth (docs/DATA_PIPELINE.md:127-134,289-291: 82,173 graph nodes,
    737,716 undirected edges, degree ~18) from 120,436 train + 23,861 val + 23,408 test sessions; the
    co-occurrence graph is built from the first `graph_sessions` (train) sessions only, as the
    reference's pipeline does.  Sessions browse one of `clusters` item groups (power-law inside the
    group) with probability p_local and the whole catalogue otherwise; that locality is what makes
    co-occurrence pairs repeat the way real sessions do.  `stats()` reports what was achieved.
    `generate_scaled()` is the 1M-item / 20M-edge configuration (BASELINE.json configs[4])."""
    rng = np.random.default_rng(seed)
    lens = session_lengths(rng, num_sessions, length_law)
    ptr = np.concatenate([[0], np.cumsum(lens)])
    total = int(ptr[-1])
    n = num_items - 1
    perm = rng.permutation(np.arange(1, num_items))

    def power_law(size, a):
        p = np.arange(1, size + 1, dtype=np.float64) ** (-a)
        return p / p.sum()

    global_pick = rng.choice(n, size=total, p=power_law(n, a_global))
    per_cluster = max(n // clusters, 1)
    sess_cluster = rng.choice(clusters, size=num_sessions, p=power_law(clusters, a_cluster))
    local_rank = rng.choice(per_cluster, size=total, p=power_law(per_cluster, a_local))
    local_pick = np.minimum(local_rank * clusters + np.repeat(sess_cluster, lens), n - 1)
    items = perm[np.where(rng.random(total) < p_local, local_pick, global_pick)]
    # revisits: with probability `revisit` an event repeats an earlier item of the same session
    pos_in_sess = np.arange(total) - np.repeat(ptr[:-1], lens)
    back = (rng.random(total) * pos_in_sess).astype(np.int64)
    redo = (rng.random(total) < revisit) & (pos_in_sess > 0)
    src_pos = np.arange(total) - 1 - back
    for _ in range(3):  # resolve short chains of revisits-of-revisits
        items = np.where(redo, items[np.maximum(src_pos, 0)], items)
    if not build_graph:   # the caller builds the co-occurrence graph on the device (data.build_co_event_graph)
        empty = np.zeros(0, dtype=np.int64)
        return SyntheticData(num_items, empty, empty, ptr, items)
    # window co-occurrence pairs over the graph (train) sessions, canonical order, counted
    upto = total if graph_sessions is None else int(ptr[min(graph_sessions, num_sessions)])
    keys = []
    for d in range(1, window + 1):
        ok = pos_in_sess[d:upto] >= d  # same session
        a, b = items[:upto - d][ok], items[d:upto][ok]
        keys.append(np.minimum(a, b) * num_items + np.maximum(a, b))
    uniq, counts = np.unique(np.concatenate(keys), return_counts=True)
    uniq = uniq[np.argsort(-counts, kind="stable")]
    return SyntheticData(num_items, uniq // num_items, uniq % num_items, ptr, items)


def build_batch(data: SyntheticData, session_ids: np.ndarray, max_len: int = 50, symmetrize: bool = False,
                self_loop_if_empty: bool = False, edge_keys: tuple | None = None) -> dict:
    """Host construction of one batch with the reference's collate semantics
    (etpgt/train/dataloader.py:84-98,126-202): context = all but the last of the (last 50) events,
    nodes = sorted unique context items, edges = stored edges with both ends in the context, in stored
    order and direction.  Pair lookup against a sorted key table instead of the reference's full-frame
    scans; integer-identical to oracle/graph_ref.collate_sessions (tests/test_synth.py)."""
    n_items = data.num_items
    if edge_keys is None:
        edge_keys = sorted_edge_keys(data)
    keys_sorted, key_order = edge_keys
    xs, srcs, dsts, batch, targets, members = [], [], [], [], [], []
    node_ptr, edge_ptr = [0], [0]
    for b, s in enumerate(session_ids):
        items = data.session(int(s))[-max_len:]
        ctx, target = items[:-1], int(items[-1])
        nodes = np.unique(ctx)
        k = len(nodes)
        iu, ju = np.triu_indices(k)
        cand = nodes[iu] * n_items + nodes[ju]
        loc = np.searchsorted(keys_sorted, cand)
        loc[loc >= len(keys_sorted)] = 0
        hit = keys_sorted[loc] == cand
        csv_idx = key_order[loc[hit]]
        stored = np.argsort(csv_idx, kind="stable")      # stored (CSV) order inside the session
        src, dst = iu[hit][stored], ju[hit][stored]
        if symmetrize:
            src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
        if self_loop_if_empty and len(src) == 0:
            src = dst = np.arange(k)
        xs.append(nodes)
        srcs.append(src + node_ptr[-1])
        dsts.append(dst + node_ptr[-1])
        batch.append(np.full(k, b, dtype=np.int64))
        targets.append(target)
        members.append(items)
        node_ptr.append(node_ptr[-1] + k)
        edge_ptr.append(edge_ptr[-1] + len(src))
    cat = lambda parts: np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, dtype=np.int64)  # noqa: E731
    return dict(x=cat(xs), edge_index=np.stack([cat(srcs), cat(dsts)]), batch=cat(batch),
                target=np.asarray(targets, dtype=np.int64), node_ptr=np.asarray(node_ptr, dtype=np.int64),
                edge_ptr=np.asarray(edge_ptr, dtype=np.int64), members=members)


def sorted_edge_keys(data: SyntheticData) -> tuple:
    keys = data.item_i * data.num_items + data.item_j
    order = np.argsort(keys, kind="stable")
    return keys[order], order


def generate_scaled(num_sessions: int = 8_000_000, num_items: int = 1_000_000, seed: int = 43,
                    build_graph: bool = False) -> SyntheticData:
    """BASELINE.json configs[4] / SURVEY.md section 8d config 5: 1,000,000 items, Yoochoose-shaped sessions
    (~8M sessions, mean length ~4, min 3) whose window-5 co-occurrence graph has ~20M undirected edges (average
    degree ~40, power-law).  By default only the sessions are generated here; the graph is built from them on the
    device (etpgt_cooc_graph_build), which is the path a 1M-item data set needs anyway."""
    return generate(num_sessions=num_sessions, graph_sessions=None, num_items=num_items, clusters=12_500, p_local=0.9,
                    a_local=0.7, a_global=0.8, a_cluster=0.6, revisit=0.25, seed=seed, length_law="yoochoose",
                    build_graph=build_graph)


def zipf_graph(num_nodes: int = 82_174, num_edges: int = 737_716, a: float = 1.5, seed: int = 44,
               max_share: float = 0.02):
    """An undirected graph whose endpoints follow the popularity law of the reference's own generator
    (`np.random.zipf(1.5, num_items)` weights, scripts/data/00_generate_synthetic_data.py:53): `num_edges` distinct
    pairs (item_i <= item_j, self pairs kept), a handful of nodes holding a large share of them — the hub-row case
    of the edge kernels.  One endpoint of every pair is drawn from the zipf weights (a node's
    share capped at `max_share`, else the heaviest draw would own almost every edge and there would not be enough
    DISTINCT pairs), the other uniformly — popular items co-occur with everything.  Returns (item_i, item_j) int64."""
    rng = np.random.default_rng(seed)
    w = rng.zipf(a, size=num_nodes - 1).astype(np.float64)
    for _ in range(8):                      # cap, renormalise, repeat: the cap holds after renormalisation
        w = np.minimum(w, w.sum() * max_share)
    p = w / w.sum()
    keys = np.zeros(0, dtype=np.int64)
    for _ in range(20):
        if len(keys) >= num_edges:
            break
        need = int((num_edges - len(keys)) * 1.25) + 1024
        x = 1 + rng.choice(num_nodes - 1, size=need, p=p)
        y = 1 + rng.integers(0, num_nodes - 1, size=need)
        keys = np.unique(np.concatenate([keys, np.minimum(x, y) * num_nodes + np.maximum(x, y)]))
    keys = rng.permutation(keys)[:num_edges]
    return keys // num_nodes, keys % num_nodes
