"""GPU parity tests of the split-bf16 tcgen05 GEMM that carries the dense node projections:
forward, both backward GEMMs (incl. deterministic split-K) and the bias gradient, against fp64."""

import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,k,n_out", [(300, 256, 1024), (128, 64, 128), (1, 256, 1024), (1000, 32, 128),
                                        (4097, 256, 256), (130, 72, 40)])
def test_linear_tensor_core_forward_backward(m, k, n_out):
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(m + k)
    x = torch.randn(m, k, generator=g, dtype=torch.float64)
    w = torch.randn(n_out, k, generator=g, dtype=torch.float64) / k ** 0.5
    b = torch.randn(n_out, generator=g, dtype=torch.float64)
    d_y = torch.randn(m, n_out, generator=g, dtype=torch.float64)
    x64, w64, b64 = (t.clone().requires_grad_(True) for t in (x, w, b))
    (x64 @ w64.t() + b64).backward(d_y)
    xc, wc, bc = (t.float().cuda().requires_grad_(True) for t in (x, w, b))
    y = ops.LinearTensorCore.apply(xc, wc, bc)
    y.backward(d_y.float().cuda())
    torch.cuda.synchronize()
    assert rel_err(y, x @ w.t() + b) < 3e-5
    assert rel_err(xc.grad, x64.grad) < 3e-5
    assert rel_err(wc.grad, w64.grad) < 3e-5
    assert rel_err(bc.grad, b64.grad) < 3e-5


def test_split_k_is_deterministic_and_accurate():
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(2)
    m, k, n_out = 20000, 256, 1024            # dW: K = 20000 nodes, 16 output tiles -> split-K
    x = torch.randn(m, k, generator=g)
    w = (torch.randn(n_out, k, generator=g) / 16).cuda().requires_grad_(True)
    d_y = torch.randn(m, n_out, generator=g).cuda()
    grads = []
    for _ in range(2):
        w.grad = None
        ops.LinearTensorCore.apply(x.cuda(), w, None).backward(d_y)
        grads.append(w.grad.clone())
    assert torch.equal(grads[0], grads[1])
    want = d_y.double().t() @ x.double().cuda()
    assert rel_err(grads[0], want) < 3e-5


def test_plain_bf16_gemm_and_split_parts():
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(4)
    x = torch.randn(200, 128, generator=g).cuda()
    w = torch.randn(256, 128, generator=g).cuda()
    hi, lo, hi_t, lo_t, ld_t, sums = ops._split(x, True, True, colsum=True)   # transposed outputs stay in the ABI
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert torch.equal(lo, (x - hi.float()).to(torch.bfloat16))
    assert torch.equal(hi_t[:, :200], hi.t()) and torch.equal(lo_t[:, :200], lo.t())
    assert rel_err(sums, x.double().sum(0)) < 1e-5
    w_hi = w.to(torch.bfloat16)
    y = ops._gemm_x3(hi, None, w_hi, None, 200, 256, 128, 128, 128, None)
    assert rel_err(y, hi.double() @ w_hi.double().t()) < 1e-5


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (304, 264, 1000), (1024, 256, 5000), (72, 40, 136)])
def test_mn_major_operands(a_mn, b_mn, m, n, k):
    """Operands stored as their transposes ([K, M] / [K, N] row-major) are read MN-major: same result
    as the K-major GEMM on explicitly transposed copies (split-bf16 x3, then plain bf16)."""
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).cuda()
    b = torch.randn(n, k, generator=g).cuda()
    want = a.double() @ b.double().t()
    a_src = a.t().contiguous() if a_mn else a          # what the caller holds
    b_src = b.t().contiguous() if b_mn else b
    a_hi, a_lo, _, _, _, _ = ops._split(a_src, True, False)
    b_hi, b_lo, _, _, _, _ = ops._split(b_src, True, False)
    got = ops._gemm_x3(a_hi, a_lo, b_hi, b_lo, m, n, k, a_src.size(1), b_src.size(1), None, a_mn=a_mn, b_mn=b_mn)
    assert rel_err(got, want) < 3e-5
    plain = ops._gemm_x3(a_hi, None, b_hi, None, m, n, k, a_src.size(1), b_src.size(1), None, a_mn=a_mn, b_mn=b_mn)
    a_r = a_hi.double().t() if a_mn else a_hi.double()
    b_r = b_hi.double().t() if b_mn else b_hi.double()
    assert rel_err(plain, a_r @ b_r.t()) < 1e-5


@pytest.mark.parametrize("n,e,dim,heads,in_dim", [(300, 1500, 256, 2, 256), (129, 400, 64, 2, 64), (77, 300, 128, 4, 72)])
def test_fused_transformer_conv_layer_matches_fp64(n, e, dim, heads, in_dim):
    """ops.TransformerConvLayer (projection GEMM + fused conv, split-gradient backward) against the fp64
    oracle conv (oracle/conv_ref.transformer_conv): output and every gradient."""
    import numpy as np

    import etpgt_b200.ops as ops
    from oracle import conv_ref

    rng = np.random.default_rng(n)
    g = torch.Generator().manual_seed(e)
    ei = torch.from_numpy(np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]))
    x = torch.randn(n, in_dim, generator=g, dtype=torch.float64)
    ws = [(torch.randn(dim, in_dim, generator=g, dtype=torch.float64) / in_dim ** 0.5) for _ in range(4)]
    bs = [torch.randn(dim, generator=g, dtype=torch.float64) * 0.1 for _ in range(4)]
    w_beta = torch.randn(1, 3 * dim, generator=g, dtype=torch.float64) * 0.2
    d_out = torch.randn(n, dim, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (x, *ws, *bs, w_beta)]
    x64, w64, b64, wb64 = leaves[0], leaves[1:5], leaves[5:9], leaves[9]
    ref = conv_ref.transformer_conv(x64, ei, w64[0], b64[0], w64[1], b64[1], w64[2], b64[2], w64[3], b64[3], wb64, heads, None)
    ref.backward(d_out)
    xc = x.float().cuda().requires_grad_(True)
    wc = torch.cat(ws).float().cuda().requires_grad_(True)      # query | key | value | skip
    bc = torch.cat(bs).float().cuda().requires_grad_(True)
    wbc = w_beta.float().cuda().requires_grad_(True)
    index = ops.GraphIndex(ei.cuda(), n)
    assert ops.fused_conv_supported(xc, in_dim, 4 * dim)
    out = ops.TransformerConvLayer.apply(xc, wc, bc, wbc, None, index, heads)
    out.backward(d_out.float().cuda())
    tol = 1e-4
    assert rel_err(out, ref) < tol
    assert rel_err(xc.grad, x64.grad) < tol
    # conv_ref argument order is (query, key, value, skip) = the fused row blocks
    assert rel_err(wc.grad, torch.cat([w.grad for w in w64])) < tol
    # the key bias is cancelled by the softmax (analytically zero): compare on the scale of the whole vector
    assert rel_err(bc.grad, torch.cat([b.grad for b in b64])) < tol
    assert rel_err(wbc.grad, wb64.grad) < tol


def _ffn_reference(x, w1, b1, w2, b2, mask1=None, mask2=None):
    """x + Linear2(mask1 * GELU(Linear1(x))) * mask2 in fp64 — etpgt/model/graph_transformer.py:109-124,163-168."""
    h = torch.nn.functional.gelu(x @ w1.t() + b1)
    if mask1 is not None:
        h = h * mask1
    y = h @ w2.t() + b2
    return x + (y if mask2 is None else y * mask2)


@pytest.mark.parametrize("n,d,f", [(3000, 256, 1024), (130, 64, 128), (1, 32, 64), (517, 64, 256)])
def test_feed_forward_block_matches_fp64(n, d, f):
    """ops.FeedForward (GEMM with the GELU epilogue -> GEMM with the residual added by the TMA reduce) and its
    backward (etpgt_gelu_bwd_split between the GEMMs) against the fp64 formula."""
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(n + f)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    w1 = torch.randn(f, d, generator=g, dtype=torch.float64) / d ** 0.5
    b1 = torch.randn(f, generator=g, dtype=torch.float64) * 0.1
    w2 = torch.randn(d, f, generator=g, dtype=torch.float64) / f ** 0.5
    b2 = torch.randn(d, generator=g, dtype=torch.float64) * 0.1
    d_out = torch.randn(n, d, generator=g, dtype=torch.float64)
    ref = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    want = _ffn_reference(*ref)
    want.backward(d_out)
    dev = [t.float().cuda().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    got = ops.FeedForward.apply(*dev, 0.0, 0.0, 0, 0)
    got.backward(d_out.float().cuda())
    torch.cuda.synchronize()
    assert rel_err(got, want) < 3e-5
    for a, b in zip(dev, ref):
        assert rel_err(a.grad, b.grad) < 5e-5


def test_feed_forward_dropout_masks_are_regenerated_in_backward():
    """With dropout the GELU epilogue and etpgt_gelu_bwd_split draw the same Philox bits: the masks are read back
    from the forward's own outputs (h == 0 where dropped) and the fp64 formula with those masks must reproduce both
    the output and every gradient; the keep rate is 1 - p."""
    import etpgt_b200.ops as ops
    from etpgt_b200.ops import call, ptr, size, stream, workspace

    n, d, f, p1, p2 = 700, 64, 256, 0.25, 0.1
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    w1 = torch.randn(f, d, generator=g, dtype=torch.float64) / d ** 0.5
    b1 = torch.randn(f, generator=g, dtype=torch.float64) * 0.1
    w2 = torch.randn(d, f, generator=g, dtype=torch.float64) / f ** 0.5
    b2 = torch.randn(d, generator=g, dtype=torch.float64) * 0.1
    d_out = torch.randn(n, d, generator=g, dtype=torch.float64)
    seed1, seed2 = 1234567, 7654321
    # the first GEMM alone: pre-activation u and the split pair of h
    xc, w1c = x.float().cuda(), w1.float().cuda()
    x_hi, x_lo, _, _, _, _ = ops._split(xc, True, False)
    w_hi, w_lo, _, _, _, _ = ops._split(w1c, True, False)
    u = torch.empty(n, f, device="cuda")
    h_hi = torch.empty(n, f, dtype=torch.bfloat16, device="cuda")
    h_lo = torch.empty_like(h_hi)
    ws = workspace(size("etpgt_gemm_bf16x3_workspace_bytes", n, f, d, 1), xc.device)
    call("etpgt_gemm_bf16x3_gelu", ptr(x_hi), ptr(x_lo), ptr(w_hi), ptr(w_lo), n, f, d, d, d, ptr(b1.float().cuda()),
         ptr(u), f, ptr(h_hi), ptr(h_lo), f, p1, seed1, ptr(ws), ws.numel(), stream())
    torch.cuda.synchronize()
    assert rel_err(u, x @ w1.t() + b1) < 3e-5
    h = h_hi.double().cpu() + h_lo.double().cpu()
    act = torch.nn.functional.gelu(x @ w1.t() + b1)
    kept = h != 0
    assert abs(kept.double().mean().item() - (1 - p1)) < 0.01
    assert rel_err(h[kept], (act / (1 - p1))[kept]) < 3e-5
    mask1 = kept.double() / (1 - p1)
    mask2 = ops.dropout_mask(n * d, p2, xc.device, seed=seed2).view(n, d).double().cpu()
    ref = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    want = _ffn_reference(*ref, mask1=mask1, mask2=mask2)
    want.backward(d_out)
    dev = [t.float().cuda().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    got = ops.FeedForward.apply(*dev, p1, p2, seed1, seed2)
    got.backward(d_out.float().cuda())
    torch.cuda.synchronize()
    assert rel_err(got, want) < 3e-5
    for a, b in zip(dev, ref):
        assert rel_err(a.grad, b.grad) < 5e-5
    again = ops.FeedForward.apply(*[t.detach() for t in dev], p1, p2, seed1, seed2)
    assert torch.equal(again, got)
