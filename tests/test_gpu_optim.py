"""GPU parity of the optimizer step (etpgt_adam_step, §8 a11/f1) against the fp64 oracle
(oracle/optim_ref.py, itself pinned to torch.optim in tests/test_optim_oracle.py) and against
torch.optim.AdamW / Adam in fp32 on the same device; gradient-sink behaviour of etpgt_b200.optim."""

import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-6   # fp32: relative to the largest parameter magnitude (a handful of roundings per step)


def _params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) for s in shapes]


@pytest.mark.parametrize("kind,wd", [("adamw", 1e-5), ("adamw", 1e-2), ("adam", 0.0), ("adam", 1e-3)])
def test_adam_step_matches_oracle_and_torch(kind, wd):
    from etpgt_b200 import optim
    from oracle import optim_ref

    # 4,096-element chunks: sizes below / at / above a chunk, unaligned tails, a 2-D table
    shapes = [(1000, 256), (4096,), (4097,), (5,), (1, 768), (3, 1365)]
    init = _params(shapes, 1)
    mine = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    ref = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    cls_m = optim.AdamW if kind == "adamw" else optim.Adam
    cls_t = torch.optim.AdamW if kind == "adamw" else torch.optim.Adam
    opt_m = cls_m(mine, lr=1e-3, weight_decay=wd, grad_sinks=False)
    opt_t = cls_t(ref, lr=1e-3, weight_decay=wd)
    oracle = [(p.double().numpy(), np.zeros(p.shape), np.zeros(p.shape)) for p in init]
    for step in range(1, 5):
        grads = [g * 0.1 ** step for g in _params(shapes, 10 + step)]
        for p, q, g in zip(mine, ref, grads):
            p.grad, q.grad = g.clone().cuda(), g.clone().cuda()
        opt_m.step()
        opt_t.step()
        oracle = [optim_ref.adam_step(p, g.double().numpy(), m, v, step, lr=1e-3, weight_decay=wd,
                                      decoupled=kind == "adamw") for (p, m, v), g in zip(oracle, grads)]
        for p, q, (want, m, v) in zip(mine, ref, oracle):
            scale = np.abs(want).max()
            assert np.abs(p.detach().double().cpu().numpy() - want).max() <= TOL * scale, (kind, step, tuple(p.shape))
            assert (p.detach() - q.detach()).abs().max().item() <= TOL * scale
            st = opt_m.state[p]
            assert np.abs(st["exp_avg"].double().cpu().numpy() - m).max() <= TOL * max(np.abs(m).max(), 1e-30)
            assert np.abs(st["exp_avg_sq"].double().cpu().numpy() - v).max() <= TOL * max(np.abs(v).max(), 1e-30)
    assert int(opt_m.state[mine[0]]["step"].item()) == 4
    # state dicts are interchangeable with torch's (checkpoint compatibility, trainer.py:175-190)
    opt_t.load_state_dict(opt_m.state_dict())
    opt_m.load_state_dict(opt_t.state_dict())


def test_adam_step_argument_checks():
    from etpgt_b200 import _lib, optim

    with pytest.raises(NotImplementedError):
        optim.AdamW([torch.nn.Parameter(torch.zeros(4, device="cuda"))], amsgrad=True)
    with pytest.raises(ValueError):
        optim.AdamW([torch.nn.Parameter(torch.zeros(4, device="cuda"))], lr=-1.0)
    arr = (optim._AdamTensor * 1)(optim._AdamTensor(0, 0, 0, 0, 8))
    with pytest.raises(RuntimeError, match="NULL"):
        _lib.call("etpgt_adam_step", arr, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1, 0, _lib.stream())
    p = torch.nn.Parameter(torch.zeros(8, device="cuda"))
    arr = (optim._AdamTensor * 1)(optim._AdamTensor(p.data_ptr(), p.data_ptr(), p.data_ptr(), p.data_ptr(), 8))
    with pytest.raises(RuntimeError, match="step"):
        _lib.call("etpgt_adam_step", arr, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 0, 0, _lib.stream())
    cpu = torch.nn.Parameter(torch.zeros(8))
    cpu.grad = torch.zeros(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        optim.AdamW([cpu]).step()


def test_gradient_sink_training_step_equals_plain_autograd():
    """The persistent table-gradient buffer (rows accumulated by the loss and embedding backward,
    cleared by the step kernel) gives the same parameters as ordinary autograd + the same optimizer."""
    from etpgt_b200 import data, optim, synth
    from etpgt_b200.model import create_graph_transformer_optimized

    d = synth.generate(num_sessions=600, graph_sessions=450, num_items=20000, clusters=16, seed=5)
    graph = data.ItemGraph(d.item_i, d.item_j, d.num_items)
    store = data.SessionStore(d.sess_ptr, d.sess_items)

    def run(sinks):
        torch.manual_seed(1)
        model = create_graph_transformer_optimized(d.num_items, 256, 256, dropout=0.0, use_laplacian_pe=False).cuda()
        opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5, grad_sinks=sinks)
        assert (len(opt._sinks) == 1) == sinks          # 20,000 x 256 x 4 B = 20 MB >= 4 MB
        model.train()
        for step in range(3):
            ids = np.arange(step * 128, (step + 1) * 128)
            batch = data.build_batch(graph, store, ids, 50, False, False)
            neg = data.sample_negatives(store, ids, d.num_items, 5, seed=3, step=step)
            loss = model.compute_loss(model(batch), batch.target_item, neg)
            opt.zero_grad()
            loss.backward()
            if sinks:   # the buffer IS .grad, holds this step's rows, and row 0 (padding) stays zero
                g = model.item_embedding.weight.grad
                assert g is opt._sinks[0].grad and g.abs().sum().item() > 0 and g[0].abs().sum().item() == 0
            opt.step()
            if sinks:
                assert model.item_embedding.weight.grad.abs().sum().item() == 0   # cleared by the kernel
        return {k: v.detach().clone() for k, v in model.state_dict().items()}, loss.item()

    a, la = run(True)
    b, lb = run(False)
    assert la == lb
    for k in a:
        assert torch.equal(a[k], b[k]), k     # same kernels, same accumulation order: bit-identical


def test_gradient_sink_zero_grad_discards_unstepped_gradient():
    from etpgt_b200 import ops, optim

    table = torch.nn.Parameter(torch.randn(8192, 256, device="cuda"))
    opt = optim.AdamW([table], grad_sinks=True)
    ids = torch.arange(1, 100, device="cuda")
    out = ops.EmbedPE.apply(ids, table, None, False, None, None, 0)
    out.sum().backward()
    assert table.grad.abs().sum().item() == 99 * 256
    out = ops.EmbedPE.apply(ids, table, None, False, None, None, 0)
    out.sum().backward()                      # accumulation across two backward calls, torch semantics
    assert table.grad.abs().sum().item() == 2 * 99 * 256
    opt.zero_grad()
    assert table.grad is not None and table.grad.abs().sum().item() == 0


def test_step_after_load_state_dict_uses_the_loaded_moments():
    """The cached launch plan holds raw pointers of the moment tensors; `load_state_dict` replaces those tensors
    (resume, restore-best) while parameter and gradient pointers stay the same.  The step after a load must read
    and update the LOADED moments (not the freed old buffers): compared against torch.optim.AdamW doing the same."""
    from etpgt_b200 import optim

    shapes = [(3000, 256), (4097,), (5,)]
    init = _params(shapes, 2)
    mine = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    ref = [torch.nn.Parameter(p.clone().cuda()) for p in init]
    opt_m = optim.AdamW(mine, lr=1e-3, weight_decay=1e-5)
    opt_t = torch.optim.AdamW(ref, lr=1e-3, weight_decay=1e-5)

    def set_grads(seed):
        for p, q, g in zip(mine, ref, _params(shapes, seed)):
            if p.grad is None:
                p.grad = g.clone().cuda()
            else:
                p.grad.copy_(g.cuda())          # same gradient storage every step, as the sink / flat views are
            q.grad = g.clone().cuda()

    for seed in (20, 21):
        set_grads(seed)
        opt_m.step()
        opt_t.step()
    saved = copy.deepcopy(opt_t.state_dict())    # state_dict() returns references to the live state tensors
    for seed in (22, 23):                       # both move on ...
        set_grads(seed)
        opt_m.step()
        opt_t.step()
    opt_m.load_state_dict(copy.deepcopy(saved))   # ... and are put back to the state after step 2
    opt_t.load_state_dict(copy.deepcopy(saved))
    for p, q in zip(mine, ref):
        p.data.copy_(q.data)
    old = [opt_m.state[p]["exp_avg"] for p in mine]
    set_grads(24)
    opt_m.step()
    opt_t.step()
    for p, q, o in zip(mine, ref, old):
        scale = q.detach().abs().max().item()
        assert (p.detach() - q.detach()).abs().max().item() <= TOL * scale
        st_m, st_t = opt_m.state[p], opt_t.state[q]
        assert st_m["exp_avg"] is o              # the loaded tensor is the one that was updated
        assert (st_m["exp_avg"] - st_t["exp_avg"]).abs().max().item() <= TOL * st_t["exp_avg"].abs().max().item()
        assert int(st_m["step"].item()) == 3


def test_gradient_sink_survives_module_zero_grad():
    """`model.zero_grad()` (set_to_none=True) detaches the persistent buffer from `.grad`; the next backward must
    find it attached again and start from zero — otherwise the table would silently stop training."""
    from etpgt_b200 import ops, optim

    table = torch.nn.Parameter(torch.randn(8192, 256, device="cuda"))
    opt = optim.AdamW([table], lr=1e-2, grad_sinks=True)
    ids = torch.arange(1, 100, device="cuda")
    ops.EmbedPE.apply(ids, table, None, False, None, None, 0).sum().backward()
    table.grad = None                          # what nn.Module.zero_grad() does: the unstepped rows are discarded
    ops.EmbedPE.apply(ids, table, None, False, None, None, 0).sum().backward()
    assert table.grad is not None and table.grad.abs().sum().item() == 99 * 256
    before = table.detach().clone()
    opt.step()
    assert (table.detach() - before)[1:100].abs().min().item() > 0      # the touched rows moved
    assert table.grad.abs().sum().item() == 0
