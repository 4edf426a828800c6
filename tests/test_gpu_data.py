"""GPU parity tests of the device data path (rows a1-a3): session subgraphs, collate layout and the
Philox negative sampler, bit-exact against oracle/graph_ref.py and against the golden fixture that
the unmodified reference SessionDataset / collate_fn produced (tests/golden/dataloader.npz)."""

import numpy as np
import pytest
import torch

from golden_util import Golden

pytestmark = pytest.mark.gpu


def _check_batch(batch, want):
    assert np.array_equal(batch.x.cpu().numpy(), want["x"])
    assert np.array_equal(batch.edge_index.cpu().numpy(), np.stack([want["edge_src"], want["edge_dst"]]))
    assert np.array_equal(batch.batch.cpu().numpy(), want["batch"])
    assert np.array_equal(batch.target_item.cpu().numpy(), want["target"])
    assert np.array_equal(batch.ptr.cpu().numpy(), want["node_ptr"])


def test_reference_dataloader_golden_is_reproduced_bit_exactly():
    from etpgt_b200 import data

    g = Golden("dataloader")
    graph = data.ItemGraph(g.raw["item_i"], g.raw["item_j"], int(g.raw["num_items"]))
    store = data.SessionStore(g.raw["sess_ptr"], g.raw["sess_items"])
    batch = data.build_batch(graph, store)
    assert np.array_equal(batch.x.cpu().numpy(), g.raw["x"])
    assert np.array_equal(batch.edge_index.cpu().numpy(), g.raw["edge_index"])
    assert np.array_equal(batch.batch.cpu().numpy(), g.raw["batch"])
    assert np.array_equal(batch.target_item.cpu().numpy(), g.raw["target"])
    assert batch.num_graphs == 40 and batch.num_nodes == len(g.raw["x"])


@pytest.mark.parametrize("symmetrize,loops", [(False, False), (True, True), (True, False), (False, True)])
def test_session_subgraphs_match_oracle(symmetrize, loops):
    from etpgt_b200 import data, synth
    from oracle import graph_ref

    d = synth.generate(num_sessions=3000, graph_sessions=2000, num_items=900, clusters=30, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    rng = np.random.default_rng(0)
    ids = rng.permutation(d.num_sessions)[:700]
    ids[:3] = np.argsort(-np.diff(d.sess_ptr))[:3]        # the longest sessions (> max_len events)
    want = graph_ref.collate_sessions([d.session(int(s)) for s in ids], d.item_i, d.item_j, 50, symmetrize, loops)
    _check_batch(data.build_batch(graph, store, ids, 50, symmetrize, loops), want)
    # the host builder used for bench setup agrees as well
    host = synth.build_batch(d, ids, 50, symmetrize, loops)
    assert np.array_equal(host["x"], want["x"])
    assert np.array_equal(host["edge_index"], np.stack([want["edge_src"], want["edge_dst"]]))


def test_session_subgraphs_edge_cases():
    from etpgt_b200 import data
    from oracle import graph_ref

    # reversed (non-canonical) stored edges, duplicate items, a one-item context, an empty batch
    item_i = np.array([5, 2, 7, 7, 3, 9])
    item_j = np.array([2, 5, 7, 3, 7, 1])
    sessions = [np.array([5, 2, 5, 2, 9]), np.array([7, 3]), np.array([3, 7, 3, 7, 3, 1]), np.array([4, 6, 8])]
    ptr = np.concatenate([[0], np.cumsum([len(s) for s in sessions])])
    graph = data.ItemGraph(item_i, item_j, 12)
    store = data.SessionStore(ptr, np.concatenate(sessions))
    for sym, loops in ((False, False), (True, True)):
        want = graph_ref.collate_sessions(sessions, item_i, item_j, 50, sym, loops)
        _check_batch(data.build_batch(graph, store, None, 50, sym, loops), want)
    empty = data.build_batch(graph, store, np.zeros(0, dtype=np.int64))
    assert empty.num_nodes == 0 and empty.num_edges == 0 and empty.num_graphs == 0
    # an edge-less graph is legal too
    none = data.ItemGraph(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), 12)
    b = data.build_batch(none, store)
    assert b.num_edges == 0 and b.num_nodes == sum(len(np.unique(s[:-1])) for s in sessions)


def test_negative_sampler_is_bit_identical_to_the_oracle_stream():
    from etpgt_b200 import data, synth
    from oracle import graph_ref

    d = synth.generate(num_sessions=500, graph_sessions=100, num_items=300, clusters=10, seed=9)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    ids = np.arange(100, 400)
    got = data.sample_negatives(store, ids, d.num_items, num_neg=5, seed=0x1234567890ABCDEF, step=7).cpu().numpy()
    for b, s in enumerate(ids):
        want = graph_ref.sample_negatives(0x1234567890ABCDEF, 7, int(s), d.session(int(s))[-50:], d.num_items, 5)
        assert np.array_equal(got[b], want), s
        assert not set(got[b]) & set(d.session(int(s))[-50:]) and got[b].min() >= 1
    # keyed by the global session index: a different sharding gives the same ids
    lo = data.sample_negatives(store, ids[:120], d.num_items, 5, 0x1234567890ABCDEF, 7).cpu().numpy()
    hi = data.sample_negatives(store, ids[120:], d.num_items, 5, 0x1234567890ABCDEF, 7).cpu().numpy()
    assert np.array_equal(np.concatenate([lo, hi]), got)
    # contiguous range addressed by session_base
    sub = data.SessionStore(d.sess_ptr[100:401] - d.sess_ptr[100], d.sess_items[d.sess_ptr[100]:d.sess_ptr[400]])
    based = data.sample_negatives(sub, None, d.num_items, 5, 0x1234567890ABCDEF, 7, session_base=100).cpu().numpy()
    assert np.array_equal(based, got)
    assert not np.array_equal(got, data.sample_negatives(store, ids, d.num_items, 5, 0x1234567890ABCDEF, 8).cpu().numpy())


def test_device_batch_trains_end_to_end():
    """Device-built batch + device negatives -> model -> loss -> backward, and the result equals the
    same step fed from the oracle-built batch."""
    from etpgt_b200 import data, synth
    from etpgt_b200.model import create_graph_transformer_optimized
    from oracle import graph_ref

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=1)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    ids = np.arange(64, 192)
    batch = data.build_batch(graph, store, ids)
    batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=0)
    torch.manual_seed(0)
    model = create_graph_transformer_optimized(d.num_items, 64, 64, dropout=0.0, use_laplacian_pe=False).cuda()
    model.train()
    loss = model.compute_loss(model(batch), batch.target_item, batch.negative_items)
    loss.backward()
    grad_a = model.item_embedding.weight.grad.clone()

    want = graph_ref.collate_sessions([d.session(int(s)) for s in ids], d.item_i, d.item_j)

    class B:
        x = torch.from_numpy(want["x"]).cuda()
        edge_index = torch.from_numpy(np.stack([want["edge_src"], want["edge_dst"]])).cuda()
        batch = torch.from_numpy(want["batch"]).cuda()

    model.zero_grad()
    for bn in model.batch_norms:
        bn.reset_running_stats()
    loss_b = model.compute_loss(model(B()), torch.from_numpy(want["target"]).cuda(), batch.negative_items)
    loss_b.backward()
    assert loss.item() == loss_b.item()
    assert torch.equal(grad_a, model.item_embedding.weight.grad)


def test_create_dataloader_reads_reference_csvs(tmp_path):
    """The reference's on-disk formats (train.csv / graph_edges.csv) -> device batches identical to what
    the reference SessionDataset + collate_fn produced for the same files (golden fixture)."""
    import csv

    from etpgt_b200.train.dataloader import create_dataloader

    g = Golden("dataloader")
    ptr = g.raw["sess_ptr"]
    with open(tmp_path / "train.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["timestamp", "visitorid", "event", "itemid", "transactionid", "session_id"])
        for s in range(len(ptr) - 1):
            for t, it in enumerate(g.raw["sess_items"][ptr[s]:ptr[s + 1]]):
                w.writerow([1000 * s + t, s, "view", int(it), "", s])
    with open(tmp_path / "graph_edges.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["item_i", "item_j", "count", "last_ts", "event_pair_hist"])
        for i, j in zip(g.raw["item_i"], g.raw["item_j"]):
            w.writerow([int(i), int(j), 1, 0, "{}"])
    loader = create_dataloader(tmp_path / "train.csv", tmp_path / "graph_edges.csv", batch_size=40, shuffle=False)
    assert loader.num_items == int(g.raw["num_items"]) and len(loader) == 1
    (batch,) = list(loader)
    assert np.array_equal(batch.x.cpu().numpy(), g.raw["x"])
    assert np.array_equal(batch.edge_index.cpu().numpy(), g.raw["edge_index"])
    assert np.array_equal(batch.batch.cpu().numpy(), g.raw["batch"])
    assert np.array_equal(batch.target_item.cpu().numpy(), g.raw["target"])
    neg = batch.negative_items.view(40, 5).cpu().numpy()          # trainer.py:87-89
    for s in range(40):
        assert not set(neg[s]) & set(g.raw["sess_items"][ptr[s]:ptr[s + 1]][-50:]) and neg[s].min() >= 1
    # smaller batches partition the same sessions, shuffled epochs differ
    small = create_dataloader(tmp_path / "train.csv", tmp_path / "graph_edges.csv", batch_size=16, shuffle=True)
    assert len(small) == 3
    first = [b.target_item.cpu() for b in small]
    second = [b.target_item.cpu() for b in small]
    assert sorted(torch.cat(first).tolist()) == sorted(g.raw["target"].tolist())
    assert not torch.equal(torch.cat(first), torch.cat(second))


@pytest.mark.parametrize("window", [1, 5, 7])
def test_co_event_graph_device_builder_is_bit_exact(window):
    """etpgt_cooc_graph_build vs oracle/graph_ref.co_event_graph (pinned to the reference function by
    tests/golden/co_event_graph.npz): identical rows in identical order."""
    from golden_util import GOLDEN
    from etpgt_b200 import data
    from oracle import graph_ref

    g = dict(np.load(GOLDEN / "co_event_graph.npz"))
    want = graph_ref.co_event_graph(g["sess_ptr"], g["sess_items"], g["timestamps"], window)
    got = data.build_co_event_graph(g["sess_ptr"], g["sess_items"], g["timestamps"], window)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)
    if window == 5:   # and therefore the reference's own edge set
        ref = {(int(a), int(b)): (int(c), int(t)) for a, b, c, t in zip(g["item_i"], g["item_j"], g["count"], g["last_ts"])}
        mine = {(int(a), int(b)): (int(c), int(t)) for a, b, c, t in zip(*[x.cpu().numpy() for x in got])}
        assert mine == ref
    no_ts = data.build_co_event_graph(g["sess_ptr"], g["sess_items"], None, window)
    assert no_ts[3] is None and all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(no_ts[:3], want[:3]))


def test_co_event_graph_edge_cases_and_synthetic_scale():
    from etpgt_b200 import data, synth
    from oracle import graph_ref

    # no sessions / single-event sessions -> no edges
    for ptr, items in (([0], []), ([0, 1, 2], [3, 4])):
        got = data.build_co_event_graph(np.asarray(ptr), np.asarray(items, dtype=np.int64), None, 5, num_items=8)
        assert got[0].numel() == 0
    # a session of one repeated item -> one self edge with all the pairs
    got = data.build_co_event_graph(np.asarray([0, 4]), np.asarray([6, 6, 6, 6]), np.asarray([5, 6, 7, 8]), 2, num_items=8)
    assert [t.tolist() for t in got] == [[6], [6], [5], [7]]
    # the synthetic RetailRocket-shaped generator builds its graph with the same rule (numpy path)
    d = synth.generate(num_sessions=3000, graph_sessions=2200, num_items=900, clusters=30, seed=2)
    ptr = d.sess_ptr[: 2200 + 1]
    items = d.sess_items[: ptr[-1]]
    i, j, c, _ = data.build_co_event_graph(ptr, items, None, 5, num_items=d.num_items)
    want = graph_ref.co_event_graph(ptr, items, None, 5)
    assert np.array_equal(i.cpu().numpy(), want[0]) and np.array_equal(j.cpu().numpy(), want[1])
    assert np.array_equal(c.cpu().numpy(), want[2])


def test_scatter_plans_give_bit_identical_table_gradients():
    """ops.prepare_batch (index + the two scatter plans, made on a side stream one step ahead) must not change
    a single bit of the step: same loss, same table gradient, same dense gradients as the inline path; the
    plan of the loss keys is found again through the trainer's `negative_items.view(B, -1)` (trainer.py:87-89),
    and a modified key tensor invalidates its plan."""
    from etpgt_b200 import data, ops, synth
    from etpgt_b200.model import create_graph_transformer_optimized

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    ids = np.arange(100, 420)
    torch.manual_seed(0)
    model = create_graph_transformer_optimized(d.num_items, 64, 64, dropout=0.0, laplacian_k=8).cuda()
    model.laplacian_pe._cached_pe = torch.randn(d.num_items, 8, device="cuda").abs()
    model.train()

    def run(prepare, loss_kind):
        batch = data.build_batch(graph, store, ids)
        batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=0).reshape(-1)
        if prepare:
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                prepared = ops.prepare_batch(batch, d.num_items)
            torch.cuda.current_stream().wait_stream(side)
            assert prepared.plan_nodes.m == batch.x.numel() and prepared.plan_loss.m == len(ids) * 6
            # loss keys: [b][0] = target, [b][1 + c] = negative c
            loss_keys = torch.cat([batch.target_item.view(-1, 1), batch.negative_items.view(len(ids), -1)], 1)
            loss_keys = loss_keys.cpu().numpy().reshape(-1)
            assert np.array_equal(prepared.plan_loss.sorted_key.cpu().numpy(), np.sort(loss_keys, kind="stable"))
            perm = prepared.plan_loss.perm.cpu().numpy()    # stable: equal keys keep ascending position
            assert np.array_equal(perm, np.argsort(loss_keys, kind="stable"))
            node_keys = batch.x.cpu().numpy()
            assert np.array_equal(prepared.plan_nodes.perm.cpu().numpy(), np.argsort(node_keys, kind="stable"))
        model.zero_grad()
        for bn in model.batch_norms:
            bn.reset_running_stats()
        launches = ops.launch_count()
        sess = model(batch)
        loss = ops.sampled_loss(sess, model.item_embedding, batch.target_item,
                                batch.negative_items.view(len(ids), -1), loss_kind)[0]
        loss.backward()
        grads = {k: p.grad.clone() for k, p in model.named_parameters()}
        return loss.item(), grads, ops.launch_count() - launches

    for kind in ("bpr", "dual"):
        loss_a, grads_a, launches_a = run(False, kind)
        loss_b, grads_b, launches_b = run(True, kind)
        assert loss_a == loss_b
        for k in grads_a:
            assert torch.equal(grads_a[k], grads_b[k]), k
        # inline path: CSR build + two scatter sorts inside the step; prepared path: none of them
        assert launches_b < launches_a - 8, (launches_a, launches_b)

    # a key tensor modified after planning must not be served by the stale plan
    batch = data.build_batch(graph, store, ids)
    batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=0)
    ops.prepare_batch(batch, d.num_items)
    assert ops._find_plan(batch.target_item, batch.negative_items) is not None
    assert ops._find_plan(batch.x) is not None
    batch.negative_items[0, 0] = 1
    assert ops._find_plan(batch.target_item, batch.negative_items) is None
    x_key = ops._plan_key(batch.x)
    del batch
    assert x_key not in ops._PLANS          # plans die with their key tensors


@pytest.mark.parametrize("with_loss,edges", [(True, True), (False, True), (True, False)])
def test_batch_prepare_equals_the_separate_builds(with_loss, edges):
    """etpgt_batch_prepare (one segmented sort for the destination order and both scatter plans) must produce
    exactly what etpgt_csr_from_coo + etpgt_scatter_plan + etpgt_scatter_plan_loss produce one by one."""
    from etpgt_b200 import ops

    g = torch.Generator().manual_seed(4)
    n, e, b, items = 3001, (7919 if edges else 0), 257, 70000      # item ids need more bits than node ids
    edge_index = torch.randint(0, n, (2, e), generator=g).cuda()
    edge_index[1, : e // 3] = 5                                        # a hub destination, many ties

    class Batch:
        pass

    batch = Batch()
    batch.x = torch.randint(1, items, (n,), generator=g).cuda()
    batch.edge_index = edge_index
    if with_loss:
        batch.target_item = torch.randint(1, items, (b,), generator=g).cuda()
        batch.negative_items = torch.randint(1, items, (b * 5,), generator=g).cuda()
    prepared = ops.prepare_batch(batch, items)
    index = ops.GraphIndex(edge_index, n)
    for name in ("rowptr", "col", "eperm", "colptr", "row", "cpos"):
        assert torch.equal(getattr(prepared.index, name), getattr(index, name)), name
    nodes = ops.ScatterPlan(batch.x, items)
    assert torch.equal(prepared.plan_nodes.sorted_key, nodes.sorted_key)
    assert torch.equal(prepared.plan_nodes.perm, nodes.perm)
    if with_loss:
        loss = ops.ScatterPlan(batch.target_item, items, negatives=batch.negative_items)
        assert torch.equal(prepared.plan_loss.sorted_key, loss.sorted_key)
        assert torch.equal(prepared.plan_loss.perm, loss.perm)
        keys = torch.cat([batch.target_item.view(-1, 1), batch.negative_items.view(b, 5)], 1).reshape(-1)
        want = torch.sort(keys, stable=True)
        assert torch.equal(loss.sorted_key.long(), want.values) and torch.equal(loss.perm.long(), want.indices)
    else:
        assert prepared.plan_loss is None
