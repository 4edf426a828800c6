"""The C++ step driver (etpgt_gt_step_run, §8 a11) against the per-operator autograd path: same kernels, same
order, same arguments => every output must be BIT-identical — loss components, session embeddings, every
parameter gradient (incl. the item table), BatchNorm running statistics and counters — with dropout (same
torch seed), every driven readout and loss, 1-3 layers, with and without Laplacian PE, with and without
prepared scatter plans, run as one call or phase by phase.  The autograd path itself is checked against the
fp64 oracle in tests/test_gpu_models.py / test_gpu_training.py; one oracle check of the driver is repeated
here."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dim=64, layers=2, heads=2, dropout=0.1, readout="mean", pe=True, sessions=300, seed=3):
    from etpgt_b200 import data, synth
    from etpgt_b200.model import create_graph_transformer_optimized

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=seed)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    ids = np.arange(50, 50 + sessions)

    def make_batch():
        batch = data.build_batch(graph, store, ids)
        batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=0).reshape(-1)
        return batch

    torch.manual_seed(0)
    model = create_graph_transformer_optimized(d.num_items, dim, dim, num_layers=layers, num_heads=heads,
                                               dropout=dropout, readout_type=readout, use_laplacian_pe=pe,
                                               laplacian_k=8).cuda()
    if pe:
        model.laplacian_pe._cached_pe = torch.randn(d.num_items, 8, device="cuda").abs()
    model.train()
    return d, model, make_batch


def _snapshot(model):
    out = {k: p.grad.clone() for k, p in model.named_parameters()}
    for l, bn in enumerate(model.batch_norms):
        out[f"rm{l}"], out[f"rv{l}"] = bn.running_mean.clone(), bn.running_var.clone()
        out[f"nbt{l}"] = bn.num_batches_tracked.clone()
    return out


def _reset(model):
    model.zero_grad(set_to_none=True)
    for bn in model.batch_norms:
        bn.reset_running_stats()


def _autograd_step(model, batch, loss_kind, seed):
    from etpgt_b200 import ops

    _reset(model)
    torch.manual_seed(seed)
    sess = model(batch)
    losses = ops.sampled_loss(sess, model.item_embedding, batch.target_item,
                              batch.negative_items.view(batch.target_item.numel(), -1), loss_kind)
    losses[0].backward()
    return losses.detach().clone(), sess.detach().clone(), _snapshot(model)


def _driver_step(model, batch, loss_kind, seed, phase_by_phase=False):
    from etpgt_b200 import _lib
    from etpgt_b200.train.step import FusedTrainStep

    _reset(model)
    torch.manual_seed(seed)
    # phase by phase: the plain driver (the data-parallel control flow of the NCCL exchange); otherwise the default,
    # which runs batches of this size as ONE CUDA graph launch
    step = FusedTrainStep(model, loss_kind, graph=False if phase_by_phase else "auto")
    if phase_by_phase:
        # single-GPU run of the data-parallel control flow: one call per phase, no exchange in between
        calls = []
        real = _lib.call

        def spy(name, *args):
            if name == "etpgt_gt_step_run":
                desc, lo, hi, stream = args
                for p in range(lo, hi):
                    calls.append(p)
                    real(name, desc, p, p + 1, stream)
                return
            return real(name, *args)

        import etpgt_b200.train.step as step_mod
        step_mod._lib.call, saved = spy, step_mod._lib.call
        try:
            losses = step(batch)
        finally:
            step_mod._lib.call = saved
        assert calls == list(range(2 * len(model.convs) + 2))
    else:
        losses = step(batch)
    return losses.clone(), step.session_embeddings.clone(), _snapshot(model)


def _assert_identical(a, b):
    (la, sa, ga), (lb, sb, gb) = a, b
    assert torch.equal(la, lb), (la, lb)
    assert torch.equal(sa, sb)
    assert ga.keys() == gb.keys()
    for k in ga:
        assert torch.equal(ga[k], gb[k]), k


@pytest.mark.parametrize("dim,layers,heads,dropout,readout,pe,loss", [
    (64, 2, 2, 0.0, "mean", True, "bpr"),
    (64, 2, 2, 0.1, "mean", True, "dual"),
    (256, 2, 2, 0.1, "mean", True, "bpr"),
    (128, 3, 4, 0.1, "max", False, "listwise"),
    (32, 1, 1, 0.2, "last", True, "bpr"),
])
def test_driver_is_bit_identical_to_the_autograd_path(dim, layers, heads, dropout, readout, pe, loss):
    from etpgt_b200 import ops

    d, model, make_batch = _setup(dim, layers, heads, dropout, readout, pe)
    want = _autograd_step(model, make_batch(), loss, seed=11)
    launches = ops.launch_count()
    got = _driver_step(model, make_batch(), loss, seed=11)
    driver_launches = ops.launch_count() - launches
    _assert_identical(want, got)
    assert driver_launches > 30          # the library's kernels ran (nothing else can have produced this)
    # prepared batch (index + scatter plans made ahead): still the same bits
    batch = make_batch()
    ops.prepare_batch(batch, d.num_items)
    _assert_identical(want, _driver_step(model, batch, loss, seed=11))
    # the data-parallel control flow (one call per phase)
    _assert_identical(want, _driver_step(model, make_batch(), loss, seed=11, phase_by_phase=True))


def test_driver_matches_the_fp64_oracle():
    from oracle import model_ref

    d, model, make_batch = _setup(dim=64, dropout=0.0)
    batch = make_batch()
    state = {k: (v.detach().double().cpu() if v.is_floating_point() else v.cpu()) for k, v in model.state_dict().items()}
    state["laplacian_pe._cached_pe"] = model.laplacian_pe._cached_pe.double().cpu()
    losses, sess, grads = _driver_step(model, batch, "bpr", seed=0)
    want = model_ref.graph_transformer_forward(state, batch.x.cpu(), batch.edge_index.cpu(), batch.batch.cpu(),
                                               num_layers=2, num_heads=2, training=True)
    negatives = batch.negative_items.view(batch.target_item.numel(), -1).cpu()
    want_loss = model_ref.bpr_loss(want, state["item_embedding.weight"], batch.target_item.cpu(), negatives)
    assert (sess.double().cpu() - want).abs().max().item() <= 1e-4 * want.abs().max().item()
    assert abs(losses[0].item() - want_loss.item()) <= 1e-4 * abs(want_loss.item())
    assert losses[2].item() == losses[0].item()      # bpr mode: total = bpr component


def test_driver_gradient_semantics_and_training_loop():
    """.grad is set when None and accumulated otherwise (torch semantics); with the device optimizer the table
    gradient goes to its persistent buffer; a short training run equals the autograd run bit for bit."""
    from etpgt_b200 import ops, optim
    from etpgt_b200.train.step import FusedTrainStep

    d, model, make_batch = _setup(dim=64, dropout=0.0)
    batch = make_batch()
    step = FusedTrainStep(model, "bpr")
    _reset(model)
    step(batch)
    once = {k: p.grad.clone() for k, p in model.named_parameters()}
    step(batch)                                   # no zero_grad in between: gradients accumulate
    for k, p in model.named_parameters():
        assert torch.allclose(p.grad, 2 * once[k], rtol=1e-6, atol=1e-12), k

    def train(use_driver):
        torch.manual_seed(0)
        d2, m, mk = _setup(dim=64, dropout=0.1)
        opt = optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-5)
        fused = FusedTrainStep(m, "bpr") if use_driver else None
        out = []
        torch.manual_seed(5)
        for _ in range(4):
            b = mk()
            opt.zero_grad()
            if use_driver:
                loss = fused(b)[0]
            else:
                loss = ops.sampled_loss(m(b), m.item_embedding, b.target_item, b.negative_items.view(-1, 5), "bpr")[0]
                loss.backward()
            opt.step()
            out.append(loss.item())
        return out, {k: v.clone() for k, v in m.state_dict().items()}

    losses_a, state_a = train(False)
    losses_b, state_b = train(True)
    assert losses_a == losses_b
    for k in state_a:
        assert torch.equal(state_a[k], state_b[k]), k


def test_driver_rejects_what_it_does_not_cover():
    from etpgt_b200.model import create_gat, create_graph_transformer
    from etpgt_b200.train.step import FusedTrainStep

    assert not FusedTrainStep.supported(create_graph_transformer(50, 64, 64).cuda())       # FFN variant
    assert not FusedTrainStep.supported(create_gat(50, 64, 64).cuda())
    with pytest.raises(NotImplementedError):
        FusedTrainStep(create_graph_transformer(50, 64, 64, use_ffn=False, readout_type="attention").cuda())
    d, model, make_batch = _setup(dim=64)
    with pytest.raises(ValueError, match="Unknown loss type"):
        FusedTrainStep(model, "hinge")


def test_trainer_mirrors_the_reference_loop(tmp_path):
    """etpgt_b200.train.trainer.Trainer (the reference's Trainer API, trainer.py:17-251): the driver-backed
    epoch equals the per-operator epoch bit for bit; evaluate() equals the oracle's Recall / NDCG on the same
    predictions; checkpoints, history and early stopping behave as in the reference."""
    import json

    from etpgt_b200 import data, optim, synth
    from etpgt_b200.model import create_graph_transformer_optimized
    from etpgt_b200.train.losses import create_loss_function
    from etpgt_b200.train.trainer import Trainer
    from oracle import model_ref

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)

    def loader(first, count, size=64):
        out = []
        for i in range(count):
            ids = np.arange(first + i * size, first + (i + 1) * size)
            batch = data.build_batch(graph, store, ids)
            batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=i).reshape(-1)
            out.append(batch)
        return out

    def run(use_driver, out_dir, loss_type):
        torch.manual_seed(0)
        model = create_graph_transformer_optimized(d.num_items, 64, 64, dropout=0.1, laplacian_k=8).cuda()
        model.laplacian_pe._cached_pe = torch.randn(d.num_items, 8, device="cuda").abs()
        opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
        loss_fn = None if loss_type is None else create_loss_function(loss_type)
        trainer = Trainer(model, loader(0, 4), loader(600, 2), opt, output_dir=out_dir, max_epochs=3, patience=1,
                          k_values=[10, 20], loss_fn=loss_fn)
        assert (trainer._driver is not None)
        if not use_driver:
            trainer._driver = None
        torch.manual_seed(9)
        history = trainer.train()
        return trainer, history

    for loss_type in (None, "dual"):
        a, hist_a = run(True, tmp_path / f"a_{loss_type}", loss_type)
        b, hist_b = run(False, tmp_path / f"b_{loss_type}", loss_type)
        assert hist_a == hist_b and len(hist_a["train_loss"]) >= 1
        for (k, v), (_, w) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
            assert torch.equal(v, w), k
    # evaluate(): device-side counters == the oracle's metrics on the gathered predictions
    metrics = a.evaluate()
    preds, targets = [], []
    a.model.eval()
    with torch.no_grad():
        for batch in a.val_loader:
            preds.append(a.model.predict(a.model(batch), k=20).cpu())
            targets.append(batch.target_item.cpu())
    preds, targets = torch.cat(preds), torch.cat(targets)
    for k in (10, 20):
        assert metrics[f"recall@{k}"] == pytest.approx(model_ref.recall_at_k(preds[:, :k], targets, k))
        assert metrics[f"ndcg@{k}"] == pytest.approx(model_ref.ndcg_at_k(preds[:, :k], targets, k))
    # artefacts of the reference loop
    out = tmp_path / "a_dual"
    assert (out / "checkpoint_latest.pt").exists() and (out / "history.json").exists()
    saved = json.loads((out / "history.json").read_text())
    assert saved == hist_a and set(saved) == {"train_loss", "val_metrics"}
    ckpt = torch.load(out / "checkpoint_latest.pt", weights_only=False)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "best_val_metric", "history"}
    assert a.patience_counter <= a.patience and len(hist_a["val_metrics"]) == len(hist_a["train_loss"])


def test_driver_edge_cases_match_the_autograd_path():
    """Ragged and degenerate batches (the shapes the reference's collate can produce): no edge at all, a single
    session, one-node sessions, a session without edges next to dense ones, an odd number of negatives."""
    from etpgt_b200.model import create_graph_transformer_optimized

    class Batch:
        pass

    def make(sizes, edge_prob, num_neg, seed):
        rng = np.random.default_rng(seed)
        ids, src, dst, bvec, off = [], [], [], [], 0
        for s, n in enumerate(sizes):
            ids.append(np.sort(rng.choice(np.arange(1, 300), size=n, replace=False)))
            pairs = [(a, b) for a in range(n) for b in range(a, n) if rng.random() < edge_prob[s % len(edge_prob)]]
            src += [off + a for a, _ in pairs]
            dst += [off + b for _, b in pairs]
            bvec += [s] * int(n)
            off += int(n)
        b = Batch()
        b.x = torch.from_numpy(np.concatenate(ids)).cuda()
        b.edge_index = torch.tensor([src, dst], dtype=torch.long).reshape(2, -1).cuda()
        b.batch = torch.tensor(bvec, dtype=torch.long).cuda()
        b.num_graphs = len(sizes)
        b.target_item = torch.from_numpy(rng.integers(1, 300, size=len(sizes))).cuda()
        b.negative_items = torch.from_numpy(rng.integers(1, 300, size=len(sizes) * num_neg)).cuda()
        return b

    torch.manual_seed(0)
    model = create_graph_transformer_optimized(300, 64, 64, dropout=0.1, laplacian_k=8).cuda()
    model.laplacian_pe._cached_pe = torch.randn(300, 8, device="cuda").abs()
    model.train()
    cases = [
        ([3, 5, 2, 7], [0.0], 5, 1),                 # no edges anywhere
        ([6], [0.6], 5, 2),                          # one session
        ([1, 1, 4, 1], [1.0], 3, 3),                 # one-node sessions (self loops only), 3 negatives
        ([5, 4, 9, 2, 8], [0.0, 0.9], 1, 4),         # edgeless sessions between dense ones, a single negative
    ]
    for sizes, prob, num_neg, seed in cases:
        batch = make(sizes, prob, num_neg, seed)
        want = _autograd_step(model, batch, "dual", seed=21)
        _assert_identical(want, _driver_step(model, make(sizes, prob, num_neg, seed), "dual", seed=21))
        prepared = make(sizes, prob, num_neg, seed)
        from etpgt_b200 import ops
        ops.prepare_batch(prepared, 300)
        _assert_identical(want, _driver_step(model, prepared, "dual", seed=21))


@pytest.mark.parametrize("on_side_stream", [False, True])
def test_graph_launch_is_bit_identical_and_only_updates_its_graph(on_side_stream):
    """etpgt_gt_step_run_graph: the step's launches captured and run as ONE CUDA graph launch — same bits as the
    plain driver for batches of different sizes (the executable graph is updated in place, not rebuilt), from the
    legacy default stream (the step moves to a stream of its own) and from a side stream."""
    from etpgt_b200 import data, ops, synth
    from etpgt_b200.model import create_graph_transformer_optimized
    from etpgt_b200.train.step import FusedTrainStep

    d = synth.generate(num_sessions=800, graph_sessions=600, num_items=500, clusters=20, seed=3)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)
    torch.manual_seed(0)
    model = create_graph_transformer_optimized(d.num_items, 64, 64, dropout=0.1, laplacian_k=8).cuda()
    model.laplacian_pe._cached_pe = torch.randn(d.num_items, 8, device="cuda").abs()
    model.train()

    def batch_of(first, count):
        ids = np.arange(first, first + count)
        batch = data.build_batch(graph, store, ids)
        batch.negative_items = data.sample_negatives(store, ids, d.num_items, 5, seed=1, step=first)
        ops.prepare_batch(batch, d.num_items)
        return batch

    plain, graphed = FusedTrainStep(model, "dual", graph=False), FusedTrainStep(model, "dual", graph=True)
    stream = torch.cuda.Stream() if on_side_stream else torch.cuda.current_stream()
    for first, count in [(0, 32), (100, 257), (400, 32), (40, 300), (500, 2)]:
        results = []
        for step in (plain, graphed):
            _reset(model)
            torch.manual_seed(5)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                losses = step(batch_of(first, count))
            torch.cuda.synchronize()
            results.append((losses.clone(), step.session_embeddings.clone(), _snapshot(model)))
        _assert_identical(*results)
    assert graphed.graph_rebuilds() == 1 and plain.graph_rebuilds() == 0
    # "auto" graphs small batches only
    auto = FusedTrainStep(model, "bpr")
    assert auto.graph == "auto" and auto.GRAPH_MAX_NODES > 1000
