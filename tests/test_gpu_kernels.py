"""GPU parity tests, kernel by kernel: the CUDA path (through the C ABI) against the CPU oracle
on the same seeded inputs.  Integer outputs are compared bit-exactly; fp32 outputs within 1e-4
relative (BASELINE.json) of the fp64 oracle."""

import numpy as np
import pytest
import torch

from golden_util import Golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4  # BASELINE.json: fp32 logits, losses and gradients within 1e-4 relative


@pytest.fixture(scope="module")
def ops():
    import etpgt_b200.ops as ops_mod

    return ops_mod


def random_graph(rng, n, e, self_loops=True, hub=False):
    src = rng.integers(0, n, size=e)
    dst = rng.integers(0, n, size=e)
    if hub and e > 8:
        dst[: e // 2] = 0  # one destination with a very long row
        src[e // 2: e // 2 + e // 4] = 1  # one source with a very long column
    if not self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    return np.stack([src, dst]).astype(np.int64)


# ------------------------------------------------------------------------------ CSR


@pytest.mark.parametrize("n,e", [(1, 0), (5, 0), (7, 10), (50, 400), (1000, 20000), (3, 5000)])
def test_csr_from_coo_bit_exact(ops, n, e):
    from oracle import graph_ref

    rng = np.random.default_rng(n * 131 + e)
    ei = random_graph(rng, n, e)
    want = graph_ref.csr_from_coo(ei[0], ei[1], n)
    idx = ops.GraphIndex(torch.from_numpy(ei).cuda(), n)
    for name in ("rowptr", "col", "eperm", "colptr", "row", "cpos"):
        got = getattr(idx, name).cpu().numpy()
        assert np.array_equal(got, want[name]), name


def test_segment_ptr(ops):
    batch = torch.tensor([0, 0, 0, 1, 3, 3, 5], device="cuda")
    got = ops.segment_ptr(batch, 7).cpu().tolist()
    assert got == [0, 3, 4, 4, 6, 6, 7, 7]
    assert ops.segment_ptr(batch[:0], 2).cpu().tolist() == [0, 0, 0]


# ------------------------------------------------------------------------------ embedding + PE


@pytest.mark.parametrize("dim,k_pe,per_node", [(32, 16, False), (64, 8, True), (256, 16, False), (128, 0, False)])
def test_embed_pe_forward_backward(ops, dim, k_pe, per_node):
    g = torch.Generator().manual_seed(dim + k_pe)
    items, n = 97, 300
    table = torch.randn(items, dim, generator=g, dtype=torch.float64)
    table[0] = 0
    ids = torch.randint(0, items, (n,), generator=g)
    pe = w = b = None
    if k_pe:
        pe = torch.randn(n if per_node else items, k_pe, generator=g, dtype=torch.float64).abs()
        w = torch.randn(dim, k_pe, generator=g, dtype=torch.float64)
        b = torch.randn(dim, generator=g, dtype=torch.float64)
    d_out = torch.randn(n, dim, generator=g, dtype=torch.float64)
    # oracle (fp64)
    t64 = table.clone().requires_grad_(True)
    w64 = w.clone().requires_grad_(True) if k_pe else None
    b64 = b.clone().requires_grad_(True) if k_pe else None
    ref = t64[ids]
    if k_pe:
        ref = ref + (pe if per_node else pe[ids]) @ w64.t() + b64
    ref.backward(d_out)
    t64.grad[0] = 0  # padding_idx=0 row keeps a zero gradient (nn.Embedding semantics)
    # CUDA
    tc = table.float().cuda().requires_grad_(True)
    wc = w.float().cuda().requires_grad_(True) if k_pe else None
    bc = b.float().cuda().requires_grad_(True) if k_pe else None
    out = ops.EmbedPE.apply(ids.cuda(), tc, pe.float().cuda() if k_pe else None, per_node, wc, bc, 0)
    out.backward(d_out.float().cuda())
    assert rel_err(out, ref) < TOL
    assert rel_err(tc.grad, t64.grad) < TOL
    if k_pe:
        assert rel_err(wc.grad, w64.grad) < TOL
        assert rel_err(bc.grad, b64.grad) < TOL


# ------------------------------------------------------------------------------ TransformerConv


def _tconv_case(ops, n, e, dim, heads, beta=True, mask=False, hub=False, seed=0):
    from oracle import conv_ref

    rng = np.random.default_rng(seed + n + e + dim)
    g = torch.Generator().manual_seed(seed + dim + heads)
    ei = torch.from_numpy(random_graph(rng, n, e, hub=hub))
    e = ei.size(1)
    x = torch.randn(n, dim, generator=g, dtype=torch.float64)
    ws = [torch.randn(dim, dim, generator=g, dtype=torch.float64) / dim ** 0.5 for _ in range(4)]
    bs = [torch.randn(dim, generator=g, dtype=torch.float64) * 0.1 for _ in range(4)]
    w_beta = torch.randn(1, 3 * dim, generator=g, dtype=torch.float64) * 0.2 if beta else None
    amask = None
    if mask:
        amask = (torch.rand(e, heads, generator=g) > 0.3).double() / 0.7
    d_out = torch.randn(n, dim, generator=g, dtype=torch.float64)
    # oracle on the projected features (so that the GEMM is not part of this comparison)
    w_cat, b_cat = torch.cat(ws), torch.cat(bs)
    qkvs64 = (x @ w_cat.t() + b_cat).requires_grad_(True)
    eye, zero = torch.eye(dim, dtype=torch.float64), torch.zeros(dim, dtype=torch.float64)

    def via_oracle(qkvs, wb):
        # feed the already-projected blocks through identity "projections"
        q, k, v, s = qkvs.split(dim, dim=1)
        n_ = q.size(0)
        c = dim // heads
        import math

        src, dst = ei[0], ei[1]
        from oracle.conv_ref import segment_softmax, _scatter_rows

        logits = (q.view(n_, heads, c)[dst] * k.view(n_, heads, c)[src]).sum(-1) / math.sqrt(c)
        alpha = segment_softmax(logits, dst, n_)
        if amask is not None:
            alpha = alpha * amask
        agg = _scatter_rows(v.view(n_, heads, c)[src] * alpha.unsqueeze(-1), dst, n_).reshape(n_, dim)
        if wb is None:
            return agg + s
        bta = torch.sigmoid(torch.cat([agg, s, agg - s], dim=-1) @ wb.t())
        return bta * s + (1 - bta) * agg

    wb64 = w_beta.clone().requires_grad_(True) if beta else None
    ref = via_oracle(qkvs64, wb64)
    ref.backward(d_out)
    # cross-check the shortcut above against the public oracle entry point once
    full = conv_ref.transformer_conv(x, ei, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2], ws[3], bs[3], w_beta, heads, amask)
    assert rel_err(ref, full) < 1e-12

    index = ops.GraphIndex(ei.cuda(), n)
    qc = qkvs64.detach().float().cuda().requires_grad_(True)
    wbc = w_beta.float().cuda().requires_grad_(True) if beta else None
    out = ops.TransformerConvFn.apply(qc, wbc, amask.float().cuda() if mask else None, index, heads)
    out.backward(d_out.float().cuda())
    assert rel_err(out, ref) < TOL, "forward"
    assert rel_err(qc.grad, qkvs64.grad) < TOL, "d_qkvs"
    if beta:
        assert rel_err(wbc.grad, wb64.grad) < TOL, "d_w_beta"
    return qc, wbc, index, d_out, out


@pytest.mark.parametrize("dim,heads", [(32, 2), (64, 2), (256, 2), (256, 4), (128, 1), (64, 8), (256, 8), (32, 8)])
def test_tconv_forward_backward(ops, dim, heads):
    _tconv_case(ops, n=257, e=1500, dim=dim, heads=heads)


def test_tconv_edge_cases(ops):
    _tconv_case(ops, n=5, e=0, dim=64, heads=2)              # no edges at all: pure gated skip
    _tconv_case(ops, n=1, e=3, dim=32, heads=2)              # a single node with self loops
    _tconv_case(ops, n=300, e=4000, dim=256, heads=2, hub=True)   # one very long row / column
    _tconv_case(ops, n=100, e=700, dim=64, heads=2, beta=False)   # beta=False variant
    _tconv_case(ops, n=100, e=700, dim=256, heads=2, mask=True)   # injected attention-dropout mask
    _tconv_case(ops, n=33, e=200, dim=32, heads=4, mask=True)


def test_tconv_backward_is_deterministic(ops):
    qc, wbc, index, d_out, _ = _tconv_case(ops, n=400, e=6000, dim=256, heads=2, hub=True, seed=3)
    grads = []
    for _ in range(2):
        qc.grad = None
        wbc.grad = None
        out = ops.TransformerConvFn.apply(qc, wbc, None, index, 2)
        out.backward(d_out.float().cuda())
        grads.append((qc.grad.clone(), wbc.grad.clone(), out.detach().clone()))
    for a, b in zip(grads[0], grads[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,e,dim,heads,beta,hub", [(257, 1500, 256, 2, True, False), (300, 4000, 256, 2, True, True),
                                                    (100, 700, 64, 2, False, False), (33, 200, 32, 4, True, False),
                                                    (5, 0, 128, 1, True, False)])
def test_tconv_backward_split_outputs(ops, n, e, dim, heads, beta, hub):
    """etpgt_tconv_bwd_split: the bf16 hi/lo gradient rows are bit-identical to splitting the fp32
    gradient of the plain backward, and the fused column sums are the bias gradient."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    rng = np.random.default_rng(n + e)
    g = torch.Generator().manual_seed(dim + heads)
    index = ops.GraphIndex(torch.from_numpy(random_graph(rng, n, e, hub=hub)).cuda(), n)
    qkvs = torch.randn(n, 4 * dim, generator=g).cuda()
    w_beta = (torch.randn(3 * dim, generator=g) * 0.2).cuda() if beta else None
    d_out = torch.randn(n, dim, generator=g).cuda()
    f32 = dict(dtype=torch.float32, device="cuda")
    out, agg = torch.empty(n, dim, **f32), torch.empty(n, dim, **f32)
    bt, m, inv_l = torch.empty(n, **f32), torch.empty(n, heads, **f32), torch.empty(n, heads, **f32)
    call("etpgt_tconv_fwd", ptr(qkvs), n, dim, heads, ptr(index.rowptr), ptr(index.col), ptr(index.eperm),
         index.num_edges, ptr(w_beta), None, ptr(out), ptr(agg), ptr(bt), ptr(m), ptr(inv_l), stream())
    ws = workspace(size("etpgt_tconv_bwd_workspace_bytes", n, index.num_edges, dim, heads), "cuda")

    def bwd(d_qkvs, hi, lo, colsum):
        d_wb = torch.empty(3 * dim, **f32) if beta else None
        call("etpgt_tconv_bwd_split", ptr(qkvs), ptr(d_out), n, dim, heads, ptr(index.rowptr), ptr(index.col),
             ptr(index.eperm), ptr(index.colptr), ptr(index.row), ptr(index.cpos), index.num_edges, ptr(w_beta), None,
             ptr(agg), ptr(bt), ptr(m), ptr(inv_l), ptr(d_qkvs), ptr(hi), ptr(lo), ptr(colsum), ptr(d_wb), ptr(ws),
             ws.numel(), stream())
        return d_wb

    d_plain = torch.empty(n, 4 * dim, **f32)
    wb_plain = bwd(d_plain, None, None, None)
    hi = torch.empty(n, 4 * dim, dtype=torch.bfloat16, device="cuda")
    lo = torch.empty_like(hi)
    colsum = torch.full((4 * dim,), float("nan"), **f32)
    wb_split = bwd(None, hi, lo, colsum)
    want_hi = d_plain.to(torch.bfloat16)
    assert torch.equal(hi, want_hi)
    assert torch.equal(lo, (d_plain - want_hi.float()).to(torch.bfloat16))
    assert rel_err(colsum, d_plain.double().sum(0)) < 1e-5
    if beta:
        assert torch.equal(wb_plain, wb_split)
    # both at once, and run-to-run determinism of the fused sums
    d_both, colsum2 = torch.empty(n, 4 * dim, **f32), torch.empty(4 * dim, **f32)
    hi2, lo2 = torch.empty_like(hi), torch.empty_like(hi)
    bwd(d_both, hi2, lo2, colsum2)
    assert torch.equal(colsum, colsum2) and torch.equal(hi, hi2) and torch.equal(lo, lo2)
    with pytest.raises(RuntimeError, match="gradient output required"):
        bwd(None, None, None, None)


# ------------------------------------------------------------------------------ BatchNorm


@pytest.mark.parametrize("dim,training,relu,res", [(32, True, False, True), (256, True, False, True),
                                                   (64, False, False, True), (256, True, True, False),
                                                   (128, False, True, False), (1024, True, False, False)])
def test_batch_norm_rows(ops, dim, training, relu, res):
    g = torch.Generator().manual_seed(dim)
    n = 777
    x = (torch.randn(n, dim, generator=g, dtype=torch.float64) * 2 + 3)
    r = torch.randn(n, dim, generator=g, dtype=torch.float64) if res else None
    gamma = torch.rand(dim, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(dim, generator=g, dtype=torch.float64) * 0.1
    rm = torch.randn(dim, generator=g, dtype=torch.float64) * 0.1 + 3
    rv = torch.rand(dim, generator=g, dtype=torch.float64) + 3.5
    d_y = torch.randn(n, dim, generator=g, dtype=torch.float64)
    from oracle.model_ref import batch_norm_rows

    x64, r64 = x.clone().requires_grad_(True), (r.clone().requires_grad_(True) if res else None)
    g64, b64 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    bn = batch_norm_rows(x64, g64, b64, rm, rv, training)
    y = bn.y + (r64 if res else 0)
    y = torch.relu(y) if relu else y
    y.backward(d_y)

    xc = x.float().cuda().requires_grad_(True)
    rc = r.float().cuda().requires_grad_(True) if res else None
    gc, bc = gamma.float().cuda().requires_grad_(True), beta.float().cuda().requires_grad_(True)
    rmc, rvc = rm.float().cuda(), rv.float().cuda()
    out = ops.BatchNormRows.apply(xc, gc, bc, rc, rmc, rvc, training, 0.1, 1e-5, relu, False)
    out.backward(d_y.float().cuda())
    assert rel_err(out, y) < TOL
    assert rel_err(xc.grad, x64.grad) < TOL
    assert rel_err(gc.grad, g64.grad) < TOL
    assert rel_err(bc.grad, b64.grad) < TOL
    if res:
        assert rel_err(rc.grad, r64.grad) < TOL
    assert rel_err(rmc, bn.running_mean) < TOL
    assert rel_err(rvc, bn.running_var) < TOL


# ------------------------------------------------------------------------------ readout


@pytest.mark.parametrize("kind", ["mean", "max", "last", "attention"])
@pytest.mark.parametrize("dim", [32, 256])
def test_session_readout(ops, kind, dim):
    from oracle.model_ref import session_readout

    g = torch.Generator().manual_seed(dim)
    sizes = [3, 1, 7, 2, 50, 4, 1, 1, 9]
    bvec = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    n = int(bvec.numel())
    x = torch.randn(n, dim, generator=g, dtype=torch.float64)
    aw = torch.randn(1, dim, generator=g, dtype=torch.float64) * 0.3
    ab = torch.tensor([0.2], dtype=torch.float64)
    d_out = torch.randn(len(sizes), dim, generator=g, dtype=torch.float64)
    x64 = x.clone().requires_grad_(True)
    aw64, ab64 = aw.clone().requires_grad_(True), ab.clone().requires_grad_(True)
    ref = session_readout(x64, bvec, len(sizes), kind, aw64, ab64)
    ref.backward(d_out)

    xc = x.float().cuda().requires_grad_(True)
    awc, abc = aw.float().cuda().requires_grad_(True), ab.float().cuda().requires_grad_(True)
    seg = ops.segment_ptr(bvec.cuda(), len(sizes))
    scores = torch.nn.functional.linear(xc, awc, abc).squeeze(-1) if kind == "attention" else None
    out = ops.SegmentReadout.apply(xc, scores, seg, ops.READOUT_MODES[kind])
    out.backward(d_out.float().cuda())
    assert rel_err(out, ref) < TOL
    assert rel_err(xc.grad, x64.grad) < TOL
    if kind == "attention":
        assert rel_err(awc.grad, aw64.grad) < TOL
        assert abc.grad.abs().max().item() < 1e-5  # analytically zero: the softmax cancels a shared bias


def test_golden_readouts(ops):
    g = Golden("loss_readout_metrics")
    x, bvec = g.tensor("ro/x").float().cuda(), g.tensor("ro/batch").cuda()
    # fixture is 32 wide
    seg = ops.segment_ptr(bvec, 4)
    for kind in ("mean", "max", "last"):
        out = ops.SegmentReadout.apply(x, None, seg, ops.READOUT_MODES[kind])
        assert rel_err(out, g.raw[f"ro/{kind}"]) < TOL, kind
    scores = torch.nn.functional.linear(x, g.tensor("ro/att_w").float().cuda(), g.tensor("ro/att_b").float().cuda())
    out = ops.SegmentReadout.apply(x, scores.squeeze(-1), seg, 3)
    assert rel_err(out, g.raw["ro/attention"]) < TOL


# ------------------------------------------------------------------------------ losses


@pytest.mark.parametrize("kind,kw", [("bpr", {}), ("listwise", {"temperature": 0.5}),
                                     ("dual", {"alpha": 0.7, "temperature": 2.0})])
def test_sampled_losses_match_reference_golden(ops, kind, kw):
    g = Golden("loss_readout_metrics")
    sess = g.tensor("sess").float().cuda().requires_grad_(True)
    emb = torch.nn.Embedding(50, 32).cuda()
    with torch.no_grad():
        emb.weight.copy_(g.tensor("table").float())
    losses = ops.sampled_loss(sess, emb, g.tensor("target").cuda(), g.tensor("negatives").cuda(), kind, **kw)
    losses[0].backward()
    assert abs(losses[0].item() - g.raw[f"{kind}/loss"].item()) < TOL * abs(g.raw[f"{kind}/loss"].item())
    assert rel_err(sess.grad, g.raw[f"{kind}/dsess"]) < TOL
    assert rel_err(emb.weight.grad, g.raw[f"{kind}/dtable"]) < TOL
    if kind == "dual":
        assert np.allclose(losses.detach().cpu().numpy(), g.raw["dual/parts"], rtol=TOL)


@pytest.mark.parametrize("dim,batch,neg", [(256, 1000, 5), (64, 33, 1), (32, 7, 20)])
def test_sampled_loss_shapes_and_padding_row(ops, dim, batch, neg):
    from oracle import model_ref

    g = torch.Generator().manual_seed(batch)
    items = 500
    sess = torch.randn(batch, dim, generator=g, dtype=torch.float64) * 0.3
    table = torch.randn(items, dim, generator=g, dtype=torch.float64) * 0.3
    table[0] = 0
    tgt = torch.randint(0, items, (batch,), generator=g)   # id 0 may appear: its row gets no gradient
    negs = torch.randint(0, items, (batch, neg), generator=g)
    s64, t64 = sess.clone().requires_grad_(True), table.clone().requires_grad_(True)
    total, lw, bp = model_ref.dual_loss(s64, t64, tgt, negs, 0.6, 1.3)
    total.backward()
    t64.grad[0] = 0
    emb = torch.nn.Embedding(items, dim, padding_idx=0).cuda()
    with torch.no_grad():
        emb.weight.copy_(table.float())
    sc = sess.float().cuda().requires_grad_(True)
    losses = ops.sampled_loss(sc, emb, tgt.cuda(), negs.cuda(), "dual", alpha=0.6, temperature=1.3)
    losses[0].backward()
    assert np.allclose(losses.detach().cpu().numpy(), [total.item(), lw.item(), bp.item()], rtol=TOL)
    assert rel_err(sc.grad, s64.grad) < TOL
    assert rel_err(emb.weight.grad, t64.grad) < TOL
    # determinism of the sorted scatter
    g1 = emb.weight.grad.clone()
    emb.weight.grad = None
    sc.grad = None
    ops.sampled_loss(sc, emb, tgt.cuda(), negs.cuda(), "dual", alpha=0.6, temperature=1.3)[0].backward()
    assert torch.equal(g1, emb.weight.grad)


def test_unknown_kinds_raise_value_error(ops):
    with pytest.raises(KeyError):
        ops.sampled_loss(torch.zeros(1, 32, device="cuda"), torch.zeros(4, 32, device="cuda"),
                         torch.zeros(1, dtype=torch.long, device="cuda"),
                         torch.zeros(1, 1, dtype=torch.long, device="cuda"), "hinge")
    from etpgt_b200.train.losses import create_loss_function

    with pytest.raises(ValueError, match="Unknown loss type"):
        create_loss_function("hinge")


# ------------------------------------------------------------------------------ scoring / top-k


@pytest.mark.parametrize("dim,batch,items,k", [(256, 32, 5000, 20), (64, 5, 300, 10), (32, 130, 1000, 20),
                                               (128, 1, 64, 64), (256, 9, 20, 20)])
def test_score_topk_random(ops, dim, batch, items, k):
    from oracle import model_ref

    g = torch.Generator().manual_seed(items)
    sess = torch.randn(batch, dim, generator=g)
    table = torch.randn(items, dim, generator=g)
    want_v, want_i = model_ref.predict(sess.double(), table.double(), k)
    got_v, got_i = ops.score_topk(sess.cuda(), table.cuda(), k)
    assert rel_err(got_v, want_v) < TOL
    # random fp32 scores: a swap can only happen between numerically tied neighbours
    same = got_i.cpu() == want_i
    gap_ok = (want_v - torch.gather(sess.double() @ table.double().t(), 1, got_i.cpu())).abs() < 1e-4
    assert bool((same | gap_ok).all())
    assert same.float().mean() > 0.99


def test_score_topk_exact_ties_go_to_lower_id(ops):
    """Small-integer operands make every product and sum exact, so heavy ties are real ties:
    indices must match the stable-sort oracle bit for bit."""
    from oracle import model_ref

    g = torch.Generator().manual_seed(5)
    sess = torch.randint(-2, 3, (40, 64), generator=g).float()
    table = torch.randint(-1, 2, (3000, 64), generator=g).float()
    table[0] = 0
    _, want = model_ref.predict(sess.double(), table.double(), 20)
    got_v, got = ops.score_topk(sess.cuda(), table.cuda(), 20)
    assert torch.equal(got.cpu(), want)
    # id_base shifts shard-local ids
    _, got2 = ops.score_topk(sess.cuda(), table.cuda(), 20, id_base=7000)
    assert torch.equal(got2.cpu(), want + 7000)


def test_topk_merge_and_metrics(ops):
    from oracle import graph_ref, model_ref

    rng = np.random.default_rng(3)
    vals = rng.integers(0, 6, size=(17, 8 * 20)).astype(np.float32)   # many ties across shards
    ids = np.stack([rng.permutation(10_000)[: 8 * 20] for _ in range(17)]).astype(np.int64)
    want_v, want_i = graph_ref.merge_topk(vals, ids, 20)
    got_v, got_i = ops.topk_merge(torch.from_numpy(vals).cuda(), torch.from_numpy(ids).cuda(), 20)
    assert np.array_equal(got_i.cpu().numpy(), want_i)
    assert np.array_equal(got_v.cpu().numpy(), want_v)
    targets = torch.from_numpy(want_i[:, 3]).clone()
    targets[::4] = 123456  # misses
    acc = ops.topk_metrics(got_i, targets.cuda(), 10)
    hits, gain = acc.cpu().tolist()
    assert hits / 17 == pytest.approx(model_ref.recall_at_k(torch.from_numpy(want_i), targets, 10))
    assert gain / 17 == pytest.approx(model_ref.ndcg_at_k(torch.from_numpy(want_i), targets, 10))
    # the reference's known answers (tests/test_utils.py:62-93)
    g = Golden("loss_readout_metrics")
    from etpgt_b200.utils.metrics import compute_ndcg_at_k, compute_recall_at_k

    preds, tg = g.tensor("met/preds").cuda(), g.tensor("met/targets").cuda()
    assert compute_recall_at_k(preds, tg, 5) == pytest.approx(2 / 3)
    assert compute_recall_at_k(preds, tg, 2) == pytest.approx(1 / 3)
    assert compute_ndcg_at_k(preds, tg, 5) == pytest.approx(g.raw["met/ndcg5"].item(), abs=1e-6)


# ------------------------------------------------------------------------------ fused dropout in the BN epilogue


@pytest.mark.parametrize("p,relu,res", [(0.1, False, True), (0.5, True, False), (0.9, False, True)])
def test_batch_norm_fused_dropout_forward_backward(ops, p, relu, res):
    """etpgt_bn_apply_ex / _bwd_stats_ex / _bwd_apply_ex: y = dropout_p(relu?(bn(x) + residual)).  The
    Philox mask is never stored: it is recovered here from y / y(p=0) and every backward quantity is
    recomputed with it in fp64."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    n, dim, seed = 3000, 256, 123456789
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, dim, generator=g).cuda()
    r = torch.randn(n, dim, generator=g).cuda() if res else None
    gamma, bias = (torch.rand(dim, generator=g) + 0.5).cuda(), torch.randn(dim, generator=g).cuda()
    mean, invstd = x.mean(0), 1.0 / (x.var(0, unbiased=False) + 1e-5).sqrt()
    d_y = torch.randn(n, dim, generator=g).cuda()

    def apply(pp, split=False):
        y = torch.empty_like(x)
        hi = torch.empty(n, dim, dtype=torch.bfloat16, device="cuda") if split else None
        lo = torch.empty_like(hi) if split else None
        call("etpgt_bn_apply_ex", ptr(x), n, dim, ptr(mean), ptr(invstd), ptr(gamma), ptr(bias), ptr(r), int(relu),
             float(pp), seed, ptr(y), ptr(hi), ptr(lo), stream())
        return y, hi, lo

    y0, _, _ = apply(0.0)
    y, hi, lo = apply(p, split=True)
    want0 = (x.double() - mean.double()) * invstd.double() * gamma.double() + bias.double()
    if res:
        want0 = want0 + r.double()
    if relu:
        want0 = want0.clamp_min(0)
    assert rel_err(y0, want0) < 1e-5
    keep = 1.0 / (1.0 - p)
    nz = y0 != 0
    factor = torch.where(nz, y / torch.where(nz, y0, torch.ones_like(y0)), torch.full_like(y0, float("nan")))
    kept = (factor - keep).abs() < 1e-5 * keep
    dropped = factor == 0
    assert bool((kept | dropped | ~nz).all())                       # every element is 0 or y0 / (1 - p)
    frac = dropped[nz].float().mean().item()
    assert abs(frac - p) < 5 * (p * (1 - p) / nz.sum().item()) ** 0.5 + 1e-3
    y2, _, _ = apply(p)
    assert torch.equal(y, y2)                                        # same seed -> same mask
    assert torch.equal(hi, y.to(torch.bfloat16)) and torch.equal(lo, (y - hi.float()).to(torch.bfloat16))
    # backward with the regenerated mask (where y0 == 0 the mask is unobservable: only ReLU zeros, whose
    # gradient is gated off anyway)
    mask = torch.where(dropped, torch.zeros_like(y0), torch.full_like(y0, keep)).double()
    gate = (y0 > 0).double() if relu else torch.ones_like(mask)
    gtrue = d_y.double() * mask * gate
    xhat = (x.double() - mean.double()) * invstd.double()
    sums = torch.empty(2 * dim + 1, dtype=torch.float64, device="cuda")
    ws = workspace(size("etpgt_bn_workspace_bytes", n, dim), "cuda")
    call("etpgt_bn_bwd_stats_ex", ptr(x), ptr(y), ptr(d_y), n, dim, ptr(mean), ptr(invstd), int(relu), float(p), seed,
         ptr(sums), ptr(ws), ws.numel(), stream())
    assert rel_err(sums[:dim], gtrue.sum(0)) < 1e-6 and rel_err(sums[dim:2 * dim], (gtrue * xhat).sum(0)) < 1e-6
    d_x, d_res = torch.empty_like(x), torch.empty_like(x)
    d_gamma, d_bias = torch.empty(dim, device="cuda"), torch.empty(dim, device="cuda")
    call("etpgt_bn_bwd_apply_ex", ptr(x), ptr(y), ptr(d_y), n, dim, ptr(mean), ptr(invstd), ptr(gamma), int(relu), 1,
         ptr(sums), float(n), ptr(sums), float(p), seed, ptr(d_x), ptr(d_res), ptr(d_gamma), ptr(d_bias), stream())
    assert rel_err(d_res, gtrue) < 1e-6
    want_dx = gamma.double() * invstd.double() * (gtrue - gtrue.mean(0) - xhat * (gtrue * xhat).mean(0))
    assert rel_err(d_x, want_dx) < 1e-5
    assert rel_err(d_bias, gtrue.sum(0)) < 1e-5 and rel_err(d_gamma, (gtrue * xhat).sum(0)) < 1e-5


# ------------------------------------------------------------------------------ GAT node-wise kernels


@pytest.mark.parametrize("n,heads,c", [(500, 4, 256), (77, 2, 64), (1, 1, 32), (300, 8, 128)])
def test_gat_scores_and_head_mean_kernels(ops, n, heads, c):
    """etpgt_gat_scores_fwd/_bwd and etpgt_head_mean_fwd/_bwd against the PyTorch expressions of PyG GATConv
    ((h * att).sum(-1), mean over heads + bias) in fp64."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    g = torch.Generator().manual_seed(n + c)
    width = heads * c
    h = torch.randn(n, width, generator=g)
    att_s, att_d = torch.randn(width, generator=g), torch.randn(width, generator=g)
    hc, sc, dc = h.cuda(), att_s.cuda(), att_d.cuda()
    a_src, a_dst = torch.empty(n, heads, device="cuda"), torch.empty(n, heads, device="cuda")
    call("etpgt_gat_scores_fwd", ptr(hc), ptr(sc), ptr(dc), n, width, heads, ptr(a_src), ptr(a_dst), stream())
    hv = h.double().view(n, heads, c)
    assert rel_err(a_src, (hv * att_s.double().view(heads, c)).sum(-1)) < 1e-5
    assert rel_err(a_dst, (hv * att_d.double().view(heads, c)).sum(-1)) < 1e-5
    d_as, d_ad = torch.randn(n, heads, generator=g), torch.randn(n, heads, generator=g)
    d_h0 = torch.randn(n, width, generator=g)
    d_h = d_h0.clone().cuda()
    d_as_c, d_ad_c = d_as.cuda(), d_ad.cuda()          # keep the device copies alive across the call
    d_att_s, d_att_d = torch.empty(width, device="cuda"), torch.empty(width, device="cuda")
    ws = workspace(size("etpgt_gat_aux_workspace_bytes", n, width), "cuda")
    call("etpgt_gat_scores_bwd", ptr(hc), ptr(sc), ptr(dc), ptr(d_as_c), ptr(d_ad_c), n, width, heads,
         ptr(d_h), ptr(d_att_s), ptr(d_att_d), ptr(ws), ws.numel(), stream())
    want_dh = d_h0.double().view(n, heads, c) + d_as.double().unsqueeze(-1) * att_s.double().view(heads, c) \
        + d_ad.double().unsqueeze(-1) * att_d.double().view(heads, c)
    assert rel_err(d_h, want_dh.reshape(n, width)) < 1e-5
    assert rel_err(d_att_s, (d_as.double().unsqueeze(-1) * hv).sum(0).reshape(-1)) < 1e-5
    assert rel_err(d_att_d, (d_ad.double().unsqueeze(-1) * hv).sum(0).reshape(-1)) < 1e-5
    # head mean + bias
    bias = torch.randn(c, generator=g)
    out = torch.empty(n, c, device="cuda")
    bias_c = bias.cuda()
    call("etpgt_head_mean_fwd", ptr(hc), ptr(bias_c), n, heads, c, ptr(out), stream())
    assert rel_err(out, hv.mean(1) + bias.double()) < 1e-5
    d_out = torch.randn(n, c, generator=g)
    d_agg, d_bias = torch.empty(n, width, device="cuda"), torch.empty(c, device="cuda")
    d_out_c = d_out.cuda()
    call("etpgt_head_mean_bwd", ptr(d_out_c), n, heads, c, ptr(d_agg), ptr(d_bias), ptr(ws), ws.numel(), stream())
    assert rel_err(d_agg, (d_out.double() / heads).unsqueeze(1).expand(n, heads, c).reshape(n, width)) < 1e-6
    assert rel_err(d_bias, d_out.double().sum(0)) < 1e-5
