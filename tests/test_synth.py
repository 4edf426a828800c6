"""CPU tests of the host-side helpers: the synthetic RetailRocket-shaped generator, the host batch
builder used for bench setup (must be integer-identical to the oracle's restatement of the
reference collate), and the session partitioner of the data-parallel path."""

import numpy as np


def test_generator_shape_statistics():
    from etpgt_b200 import synth

    d = synth.generate(num_sessions=20000, graph_sessions=15000, num_items=9000, clusters=170, seed=1)
    s = d.stats()
    assert s["len_median"] == 4 and 5.0 < s["len_mean"] < 6.0 and np.diff(d.sess_ptr).min() == 3
    assert d.item_i.min() >= 1 and d.item_j.max() < d.num_items and bool((d.item_i <= d.item_j).all())
    keys = d.item_i * d.num_items + d.item_j
    assert len(np.unique(keys)) == len(keys)            # one row per undirected pair
    assert d.sess_items.min() >= 1                      # id 0 is the padding item


def test_host_batch_builder_matches_oracle_collate():
    from etpgt_b200 import synth
    from oracle import graph_ref

    d = synth.generate(num_sessions=1500, graph_sessions=1000, num_items=700, clusters=25, seed=2)
    ids = np.concatenate([np.argsort(-np.diff(d.sess_ptr))[:2], np.arange(100, 300)])
    for sym, loops in ((False, False), (True, True)):
        got = synth.build_batch(d, ids, 50, sym, loops)
        want = graph_ref.collate_sessions([d.session(int(s)) for s in ids], d.item_i, d.item_j, 50, sym, loops)
        assert np.array_equal(got["x"], want["x"])
        assert np.array_equal(got["edge_index"], np.stack([want["edge_src"], want["edge_dst"]]))
        assert np.array_equal(got["batch"], want["batch"])
        assert np.array_equal(got["target"], want["target"])


def test_session_partition_balances_cost():
    from etpgt_b200.parallel import item_shard, partition_sessions

    rng = np.random.default_rng(0)
    cost = rng.integers(3, 60, size=10000)
    cuts = partition_sessions(cost, 8)
    assert cuts[0] == 0 and cuts[-1] == len(cost) and bool((np.diff(cuts) > 0).all())
    loads = np.add.reduceat(cost, cuts[:-1])
    assert loads.max() / loads.mean() < 1.02
    spans = [item_shard(82174, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 82174
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
