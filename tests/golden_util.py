"""Helpers shared by the parity tests: load a tests/golden fixture written by
oracle/make_golden.py and the error measure the tolerances are stated in."""

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


class Golden:
    def __init__(self, name):
        self.raw = dict(np.load(GOLDEN / f"{name}.npz"))

    def tensor(self, key, dtype=None):
        t = torch.from_numpy(np.asarray(self.raw[key]))
        return t if dtype is None else t.to(dtype)

    def group(self, prefix, dtype=None):
        n = len(prefix) + 1
        return {k[n:]: self.tensor(k, dtype) for k in self.raw if k.startswith(prefix + "/")}

    def cfg(self):
        out = {}
        for k, v in self.raw.items():
            if k.startswith("cfg_"):
                out[k[4:]] = v.item() if hasattr(v, "item") else v
        return out


def rel_err(got, want, floor=1e-9):
    """max |got - want| / max(|want|): the 'relative' of BASELINE.json's 1e-4 bound.  `floor`
    keeps analytically-zero tensors (e.g. the key-bias gradient, which the per-destination
    softmax cancels exactly) from dividing rounding noise by rounding noise."""
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.numel() == 0:
        return 0.0
    denom = want.abs().max().item()
    return (got - want).abs().max().item() / max(denom, floor)
