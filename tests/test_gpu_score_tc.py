"""GPU parity tests of the tcgen05 / TMA scoring kernel with the fused top-k epilogue.
Oracle: fp64 scores of the bf16-rounded operands, stable descending sort (ties -> lower id)."""

import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


def _oracle(sess, table, k):
    from oracle import model_ref

    return model_ref.predict(sess.double(), table.double(), k, bf16_inputs=True)


@pytest.mark.parametrize("dim,batch,items,k", [(256, 128, 256, 20), (256, 32, 5000, 20), (64, 5, 300, 10),
                                               (128, 200, 3000, 32), (256, 300, 82174, 20), (192, 129, 1000, 1),
                                               (256, 1, 20, 20)])
def test_score_topk_tensor_core_random(dim, batch, items, k):
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(items + dim)
    sess = torch.randn(batch, dim, generator=g)
    table = torch.randn(items, dim, generator=g)
    want_v, want_i = _oracle(sess, table, k)
    got_v, got_i = ops.score_topk(sess.cuda(), table.cuda(), k, precision="bf16")
    torch.cuda.synchronize()
    assert rel_err(got_v, want_v) < 1e-4          # same bf16 operands, fp32 vs fp64 accumulation
    exact = sess.to(torch.bfloat16).double() @ table.to(torch.bfloat16).double().t()
    same = got_i.cpu() == want_i
    near_tie = (want_v - torch.gather(exact, 1, got_i.cpu().clamp(0, items - 1))).abs() < 1e-3
    assert bool((same | near_tie).all())
    assert same.float().mean() > 0.99
    # against the fp32 reference the bf16 GEMM is within 1e-2 (BASELINE.json)
    full = sess.double() @ table.double().t()
    ref_v = torch.sort(full, dim=1, descending=True).values[:, :k]
    assert rel_err(got_v, ref_v) < 1e-2


def test_score_topk_tensor_core_exact_ties():
    """Small integers are exact in bf16 and in fp32 accumulation: indices must match bit for bit,
    including the many ties (lower id wins) and the tail tile that hangs over the table end."""
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(11)
    sess = torch.randint(-2, 3, (260, 256), generator=g).float()
    table = torch.randint(-1, 2, (7001, 256), generator=g).float()
    table[0] = 0
    want_v, want_i = _oracle(sess, table, 20)
    got_v, got_i = ops.score_topk(sess.cuda(), table.cuda(), 20, precision="bf16")
    assert torch.equal(got_i.cpu(), want_i)
    assert torch.equal(got_v.cpu().double(), want_v)
    _, shifted = ops.score_topk(sess.cuda(), table.cuda(), 20, id_base=100000, precision="bf16")
    assert torch.equal(shifted.cpu(), want_i + 100000)
    # the fp32 CUDA-core scorer agrees on exact inputs
    _, f32_i = ops.score_topk(sess.cuda(), table.cuda(), 20, precision="fp32")
    assert torch.equal(f32_i.cpu(), want_i)


def test_score_topk_tensor_core_overflow_fallback(monkeypatch):
    """Adversarial score order (every chunk beats the running threshold) and a forced tiny slot buffer:
    the overflowed rows are recomputed exactly by the CUDA-core fallback kernel."""
    import etpgt_b200.ops as ops

    g = torch.Generator().manual_seed(3)
    # scores increase with the item id for the first 100 rows -> every chunk is a new maximum
    sess = torch.randint(1, 3, (150, 64), generator=g).float()
    sess[100:] = torch.randint(-2, 3, (50, 64), generator=g).float()
    table = torch.randint(0, 2, (9000, 64), generator=g).float()
    table[:, 0] = (torch.arange(9000) // 64).float()          # slowly growing column 0 (exact in bf16)
    want_v, want_i = _oracle(sess, table, 20)
    got_v, got_i = ops.score_topk(sess.cuda(), table.cuda(), 20, precision="bf16")
    assert torch.equal(got_i.cpu(), want_i) and torch.equal(got_v.cpu().double(), want_v)
    monkeypatch.setenv("ETPGT_SCORE_CAP", "2")       # every row overflows its slot buffer
    g2 = torch.Generator().manual_seed(4)
    sess2 = torch.randint(-2, 3, (70, 128), generator=g2).float()
    table2 = torch.randint(-1, 2, (5000, 128), generator=g2).float()
    want_v2, want_i2 = _oracle(sess2, table2, 20)
    got_v2, got_i2 = ops.score_topk(sess2.cuda(), table2.cuda(), 20, precision="bf16")
    assert torch.equal(got_i2.cpu(), want_i2) and torch.equal(got_v2.cpu().double(), want_v2)


def test_bf16_conversion_matches_torch():
    import etpgt_b200.ops as ops

    x = torch.randn(1000, 64) * 100
    x[0, :4] = torch.tensor([1.00390625, -1.00390625, 3.0e38, 1e-40])   # ties-to-even, large, denormal
    assert torch.equal(ops.to_bf16(x.cuda()).cpu(), x.to(torch.bfloat16))


@pytest.mark.parametrize("cap", [None, "2"])
def test_target_positions_come_out_of_the_select_epilogue(monkeypatch, cap):
    """etpgt_score_topk_bf16_eval: hit_pos[b] = position of the row's target among its k results or -1 (the input of
    Recall@k / NDCG@k, etpgt/utils/metrics.py:6-66), from the select kernel's registers — and from the CUDA-core
    fallback for rows whose slot buffer overflowed (cap = 2 forces every row there)."""
    import etpgt_b200.ops as ops
    from oracle import model_ref

    if cap is not None:
        monkeypatch.setenv("ETPGT_SCORE_CAP", cap)
    g = torch.Generator().manual_seed(9)
    batch, items, dim, k = 300, 4000, 256, 20
    sess, table = torch.randn(batch, dim, generator=g).cuda(), torch.randn(items, dim, generator=g).cuda()
    _, plain = ops.score_topk(sess, table, k, precision="bf16")
    # targets: for two thirds of the rows an id that IS in the list (positions spread over 0..k-1), else a miss
    pos = torch.arange(batch) % k
    targets = plain.cpu()[torch.arange(batch), pos].clone()
    miss = torch.arange(batch) % 3 == 0
    targets[miss] = items + 5
    val, idx, hit = ops.score_topk(sess, table, k, precision="bf16", targets=targets.cuda())
    assert torch.equal(idx, plain)
    want = torch.where(miss, torch.full_like(pos, -1), pos).int()
    assert torch.equal(hit.cpu(), want)
    for kk in (10, 20):
        acc = ops.hit_metrics(hit, kk).cpu()
        assert acc[0].item() == pytest.approx(model_ref.recall_at_k(idx.cpu()[:, :kk], targets, kk) * batch)
        assert acc[1].item() == pytest.approx(model_ref.ndcg_at_k(idx.cpu()[:, :kk], targets, kk) * batch)
