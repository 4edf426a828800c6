"""TEST INFRASTRUCTURE ONLY — CPU stand-in for the eight `torch_geometric` symbols the
reference imports.

`torch_geometric` (requirements.txt:5 of the reference, `>=2.4.0`, no lock file) is not
installed in this image and is not vendored under /root/reference, so its published
algorithms are restated here in plain torch.  Parity at this boundary is UNPINNED by the
reference (its tests assert shapes / finiteness only); structural pins are the published
parameter counts (docs/EXPERIMENTS.md:85-88), which `tests/test_oracle.py` checks.

This directory is put on `sys.path` only by `oracle/` scripts, `tests/`, `smoke()` and
`bench.py --impl reference` / `cpu_baseline`.  The product package never imports it.
"""

__version__ = "0.0-oracle-shim"

from . import data, nn, utils  # noqa: F401
