"""Hub rows of the fused TransformerConv (csrc/tconv_hub.cu): destinations / sources with more than 256 edges are
cut into chunks, one CTA per chunk with the row's query staged in shared memory, partials combined in a fixed order.

Checks: the plan against a numpy restatement (bit-exact, deterministic); forward and every gradient against the fp64
oracle (oracle/conv_ref.py, <= 1e-4 as everywhere) on graphs whose head rows hold hundreds to tens of thousands of
edges (incl. attention-dropout masks, beta=False, all lane-group geometries); against the plain row kernels on the
same inputs (the only difference is the summation order); bf16 split outputs + fused bias column sums; two runs give
the same bits."""

import numpy as np
import pytest
import torch

from golden_util import rel_err
from test_gpu_kernels import _tconv_case, random_graph

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from etpgt_b200 import ops as _ops

    return _ops


def zipf_graph(rng, n, e):
    """Endpoints drawn with the reference generator's popularity law (np.random.zipf(1.5) weights,
    scripts/data/00_generate_synthetic_data.py:53): a few nodes take a large share of the edges."""
    w = rng.zipf(1.5, size=n).astype(np.float64)
    p = w / w.sum()
    return np.stack([rng.choice(n, size=e, p=p), rng.choice(n, size=e, p=p)]).astype(np.int64)


def test_hub_plan_matches_numpy(ops):
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    rng = np.random.default_rng(5)
    n, e = 5000, 120_000
    ei = zipf_graph(rng, n, e)
    index = ops.GraphIndex(torch.from_numpy(ei).cuda(), n)
    plans = []
    for _ in range(2):
        plan = torch.zeros(size("etpgt_hub_plan_bytes", e), dtype=torch.uint8, device="cuda")
        ws = workspace(size("etpgt_hub_plan_workspace_bytes", n), "cuda")
        call("etpgt_hub_plan", ptr(index.rowptr), ptr(index.colptr), n, e, ptr(plan), ptr(ws), ws.numel(), stream())
        plans.append(plan.cpu().numpy().copy())
    assert np.array_equal(plans[0], plans[1])          # prefix sums, no atomics: the same bytes every time
    raw = plans[0]
    counts = raw[:16].view(np.int32)
    cap_rows = e // 257 + 1
    cap_chunks = e // 256 + cap_rows + 1
    pad = lambda b: (b + 255) // 256 * 256            # noqa: E731
    off = 256
    for side, ptr_t in enumerate((index.rowptr, index.colptr)):
        p = ptr_t.cpu().numpy().astype(np.int64)
        deg = np.diff(p)
        hubs = np.nonzero(deg > 256)[0]
        rows = raw[off:off + 16 * cap_rows].view(np.int32).reshape(-1, 4)
        off += pad(16 * cap_rows)
        chunks = raw[off:off + 16 * cap_chunks].view(np.int32).reshape(-1, 4)
        off += pad(16 * cap_chunks)
        nch = (deg[hubs] + 255) // 256
        assert counts[2 * side] == len(hubs) and counts[2 * side + 1] == nch.sum() and len(hubs) > 0
        first = np.concatenate([[0], np.cumsum(nch)[:-1]])
        assert np.array_equal(rows[:len(hubs), 0], hubs) and np.array_equal(rows[:len(hubs), 1], first)
        assert np.array_equal(rows[:len(hubs), 2], nch)
        c = 0
        for slot, node in enumerate(hubs):
            for k in range(nch[slot]):
                want = [node, p[node] + 256 * k, min(256, deg[node] - 256 * k), slot]
                assert chunks[c].tolist() == want
                c += 1
    # a graph without long rows has no plan at all
    small = ops.GraphIndex(torch.from_numpy(random_graph(rng, 400, 2000)).cuda(), 400)
    assert small.hub_plan() is None and index.hub_plan() is not None


@pytest.mark.parametrize("dim,heads,beta,mask", [(256, 2, True, False), (256, 8, True, True), (128, 4, False, False),
                                                 (64, 2, True, True), (32, 1, True, False)])
def test_hub_rows_match_the_fp64_oracle(ops, dim, heads, beta, mask):
    # n = 300, e = 6,000: node 0 has 3,000 in-edges (12 chunks), node 1 has 1,500 out-edges (6 chunks)
    qc, wbc, index, d_out, out = _tconv_case(ops, n=300, e=6000, dim=dim, heads=heads, beta=beta, mask=mask, hub=True)
    assert index.hub_plan() is not None


def test_zipf_graph_matches_oracle_and_the_row_kernels(ops):
    """Reference-shaped popularity: the oracle bound, and hub kernels vs the plain row kernels on the same inputs."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    rng = np.random.default_rng(11)
    n, e, dim, heads = 2000, 60_000, 256, 2
    ei = zipf_graph(rng, n, e)
    index = ops.GraphIndex(torch.from_numpy(ei).cuda(), n)
    plan = index.hub_plan()
    counts = plan[:16].view(torch.int32).tolist()
    deg_in = np.bincount(ei[1], minlength=n)
    assert counts[0] >= 3 and deg_in.max() > 5000
    g = torch.Generator().manual_seed(4)
    qkvs = (torch.randn(n, 4 * dim, generator=g) * 0.5).cuda()
    w_beta = (torch.randn(3 * dim, generator=g) * 0.2).cuda()
    d_out = torch.randn(n, dim, generator=g).cuda()
    f32 = dict(dtype=torch.float32, device="cuda")
    ws = workspace(size("etpgt_tconv_bwd_workspace_bytes", n, e, dim, heads), "cuda")
    hub_ws = workspace(size("etpgt_tconv_hub_workspace_bytes", e, dim), "cuda")

    def run(use_hubs):
        out, agg = torch.empty(n, dim, **f32), torch.empty(n, dim, **f32)
        bt, m, inv_l = torch.empty(n, **f32), torch.empty(n, heads, **f32), torch.empty(n, heads, **f32)
        call("etpgt_tconv_fwd_hub", ptr(qkvs), n, dim, heads, ptr(index.rowptr), ptr(index.col), ptr(index.eperm), e,
             ptr(w_beta), None, ptr(out), ptr(agg), ptr(bt), ptr(m), ptr(inv_l), ptr(plan) if use_hubs else None,
             ptr(hub_ws) if use_hubs else None, hub_ws.numel() if use_hubs else 0, stream())
        hi = torch.empty(n, 4 * dim, dtype=torch.bfloat16, device="cuda")
        lo = torch.empty_like(hi)
        d_qkvs, colsum, d_wb = torch.empty(n, 4 * dim, **f32), torch.empty(4 * dim, **f32), torch.empty(3 * dim, **f32)
        # gradient rows come out either as fp32 or as the bf16 hi / lo pair: one call each
        for fp32_rows in (True, False):
            call("etpgt_tconv_bwd_split_hub", ptr(qkvs), ptr(d_out), n, dim, heads, ptr(index.rowptr), ptr(index.col),
                 ptr(index.eperm), ptr(index.colptr), ptr(index.row), ptr(index.cpos), e, ptr(w_beta), None, ptr(agg),
                 ptr(bt), ptr(m), ptr(inv_l), ptr(d_qkvs) if fp32_rows else None, None if fp32_rows else ptr(hi),
                 None if fp32_rows else ptr(lo), ptr(colsum), ptr(d_wb), ptr(ws), ws.numel(),
                 ptr(plan) if use_hubs else None, ptr(hub_ws) if use_hubs else None,
                 hub_ws.numel() if use_hubs else 0, stream())
        return dict(out=out, agg=agg, beta=bt, m=m, inv_l=inv_l, d_qkvs=d_qkvs, hi=hi, lo=lo, colsum=colsum, d_wb=d_wb)

    a, b, again = run(True), run(False), run(True)
    for k in a:
        assert torch.equal(a[k], again[k]), f"{k}: two hub runs differ"         # bit-reproducible
    assert torch.equal(a["m"], b["m"])                                           # the row maximum is exact
    # the split outputs are the split of the hub path's own fp32 gradient, the column sums its bias gradient
    want_hi = a["d_qkvs"].to(torch.bfloat16)
    assert torch.equal(a["hi"], want_hi) and torch.equal(a["lo"], (a["d_qkvs"] - want_hi.float()).to(torch.bfloat16))
    assert rel_err(a["colsum"], a["d_qkvs"].double().sum(0)) < 1e-5
    # fp64 oracle (forward and, through autograd, every gradient) on the same projected features
    import math

    from oracle.conv_ref import _scatter_rows, segment_softmax

    q64 = qkvs.double().cpu().requires_grad_(True)
    wb64 = w_beta.double().cpu().view(1, -1).requires_grad_(True)
    q, k, v, s = q64.split(dim, dim=1)
    c = dim // heads
    src, dst = torch.from_numpy(ei[0]), torch.from_numpy(ei[1])
    logits = (q.view(n, heads, c)[dst] * k.view(n, heads, c)[src]).sum(-1) / math.sqrt(c)
    alpha = segment_softmax(logits, dst, n)
    agg = _scatter_rows(v.view(n, heads, c)[src] * alpha.unsqueeze(-1), dst, n).reshape(n, dim)
    bta = torch.sigmoid(torch.cat([agg, s, agg - s], dim=-1) @ wb64.t())
    ref = bta * s + (1 - bta) * agg
    ref.backward(d_out.double().cpu())
    want = {"out": ref.detach(), "agg": agg.detach(), "d_qkvs": q64.grad, "d_wb": wb64.grad.view(-1)}
    for name, w in want.items():
        assert rel_err(a[name], w) < TOL, f"hub kernels, {name}"
        # rows of > 10,000 edges through ONE lane group's serial fp32 walk (the plain row kernels): the same
        # quantities, one order of magnitude looser — the chunked hub path is also the more accurate one
        assert rel_err(b[name], w) < 10 * TOL, f"row kernels, {name}"


@pytest.mark.parametrize("graph_kind,dim,heads", [("sessions", 256, 2), ("zipf", 256, 2), ("sessions", 64, 4), ("zipf", 32, 1)])
def test_forward_with_fused_batchnorm_statistics(ops, graph_kind, dim, heads):
    """etpgt_tconv_fwd_bn: the persistent forward that also sums out and out^2 per column (double) — outputs
    bit-identical to the plain forward, statistics equal to a separate etpgt_bn_stats pass over `out` (both are
    exact-order double sums of the same floats: <= 1e-13 relative), identical bits on a second run."""
    from etpgt_b200._lib import call, ptr, size, stream, workspace

    rng = np.random.default_rng(dim + heads)
    if graph_kind == "zipf":
        n, e = 3000, 90_000
        ei = zipf_graph(rng, n, e)
    else:   # many tiny components, like a session batch
        n, e = 20_000, 58_000
        src = rng.integers(0, n, size=e)
        ei = np.stack([src, np.minimum(src + rng.integers(0, 4, size=e), n - 1)]).astype(np.int64)
    index = ops.GraphIndex(torch.from_numpy(ei).cuda(), n)
    plan = index.hub_plan()
    assert (plan is not None) == (graph_kind == "zipf")
    g = torch.Generator().manual_seed(dim)
    qkvs = (torch.randn(n, 4 * dim, generator=g) * 0.5).cuda()
    w_beta = (torch.randn(3 * dim, generator=g) * 0.2).cuda()
    f32 = dict(dtype=torch.float32, device="cuda")
    hub_ws = workspace(size("etpgt_tconv_hub_workspace_bytes", e, dim), "cuda")
    bn_ws = workspace(size("etpgt_tconv_fwd_bn_workspace_bytes", dim), "cuda")

    def run(with_stats):
        out, agg = torch.empty(n, dim, **f32), torch.empty(n, dim, **f32)
        bt, m, inv_l = torch.empty(n, **f32), torch.empty(n, heads, **f32), torch.empty(n, heads, **f32)
        sums = torch.full((2 * dim + 1,), float("nan"), dtype=torch.float64, device="cuda")
        common = (ptr(qkvs), n, dim, heads, ptr(index.rowptr), ptr(index.col), ptr(index.eperm), e, ptr(w_beta), None,
                  ptr(out), ptr(agg), ptr(bt), ptr(m), ptr(inv_l), ptr(plan), ptr(hub_ws) if plan is not None else None,
                  hub_ws.numel() if plan is not None else 0)
        if with_stats:
            call("etpgt_tconv_fwd_bn", *common, ptr(sums), ptr(bn_ws), bn_ws.numel(), stream())
        else:
            call("etpgt_tconv_fwd_hub", *common, stream())
        return out, agg, bt, m, inv_l, sums

    plain, fused, again = run(False), run(True), run(True)
    for a, b in zip(plain[:5], fused[:5]):
        assert torch.equal(a, b)
    assert torch.equal(fused[5][:2 * dim], again[5][:2 * dim])
    want = torch.empty(2 * dim, dtype=torch.float64, device="cuda")
    ws = workspace(size("etpgt_bn_workspace_bytes", n, dim), "cuda")
    call("etpgt_bn_stats", ptr(fused[0]), n, dim, ptr(want), ptr(ws), ws.numel(), stream())
    assert rel_err(fused[5][:2 * dim], want) < 1e-13
    exact = torch.cat([fused[0].double().sum(0), (fused[0].double() ** 2).sum(0)])
    assert rel_err(fused[5][:2 * dim], exact) < 1e-12
