"""CPU: the optimizer oracle (oracle/optim_ref.py) against torch.optim.AdamW / Adam — the optimizer
classes the reference itself instantiates (scripts/train/train_baseline.py:252-256,
scripts/pipeline/run_full_pipeline.py:210)."""

import numpy as np
import pytest
import torch

from oracle import optim_ref


@pytest.mark.parametrize("kind,wd", [("adamw", 1e-5), ("adamw", 1e-2), ("adam", 0.0), ("adam", 1e-3)])
def test_optimizer_oracle_matches_torch(kind, wd):
    g = torch.Generator().manual_seed(3)
    shapes = [(37, 16), (5,), (1, 48)]
    params = [torch.randn(*s, generator=g, dtype=torch.float64).requires_grad_(True) for s in shapes]
    cls = torch.optim.AdamW if kind == "adamw" else torch.optim.Adam
    opt = cls(params, lr=1e-3, weight_decay=wd)
    mine = [(p.detach().numpy().copy(), np.zeros(p.shape), np.zeros(p.shape)) for p in params]
    for step in range(1, 6):
        grads = [torch.randn(*s, generator=g, dtype=torch.float64) * 10.0 ** (-step) for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        opt.step()
        mine = [optim_ref.adam_step(p, gr.numpy(), m, v, step, lr=1e-3, weight_decay=wd, decoupled=kind == "adamw")
                for (p, m, v), gr in zip(mine, grads)]
        for p, (q, m, v) in zip(params, mine):
            assert np.abs(p.detach().numpy() - q).max() <= 1e-12 * max(1.0, np.abs(q).max())
            st = opt.state[p]
            assert np.abs(st["exp_avg"].numpy() - m).max() <= 1e-13
            assert np.abs(st["exp_avg_sq"].numpy() - v).max() <= 1e-13
