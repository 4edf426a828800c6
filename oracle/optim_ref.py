"""TEST INFRASTRUCTURE — CPU oracle of the optimizer step (not part of the product path).

Restates the dense single-tensor update that `torch.optim.AdamW` / `torch.optim.Adam` perform for
the reference (scripts/train/train_baseline.py:252-256 `AdamW(lr, weight_decay)`;
scripts/pipeline/run_full_pipeline.py:210 `Adam(lr=0.001)`; stepped at etpgt/train/trainer.py:125-127)
in numpy float64, in torch's operation order (torch/optim/adamw.py `_single_tensor_adam`):

    AdamW: p *= 1 - lr*wd          Adam: g = g + wd*p
    m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)

Pinned in tests/test_oracle.py against torch.optim.AdamW / Adam themselves (the reference's
optimizer is torch's, which IS importable here), so parity is pinned for this row.
"""

from __future__ import annotations

import numpy as np


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              decoupled=True):
    """One update; returns new (param, exp_avg, exp_avg_sq) as float64 arrays.  `step` is 1-based."""
    p = np.asarray(param, dtype=np.float64).copy()
    g = np.asarray(grad, dtype=np.float64).copy()
    m = np.asarray(exp_avg, dtype=np.float64).copy()
    v = np.asarray(exp_avg_sq, dtype=np.float64).copy()
    b1, b2 = betas
    if decoupled:
        p *= 1.0 - lr * weight_decay
    elif weight_decay != 0.0:
        g = g + weight_decay * p
    m = m + (1.0 - b1) * (g - m)
    v = b2 * v + (1.0 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
