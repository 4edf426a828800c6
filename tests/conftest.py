"""pytest configuration: registers the `gpu` marker and puts the repo root (oracle/, the
package directory) on sys.path.  Tests marked `gpu` are the CUDA parity tests proper and are
skipped automatically when no device is visible."""

import os
import sys
from pathlib import Path

# tests/test_gpu_peer.py runs several data-parallel ranks inside ONE process: while rank 0's exchange kernel spins
# on the device waiting for rank 1, the host must stay free to launch rank 1's kernels.  With lazy module loading
# the first launch of a not-yet-loaded kernel can wait for the device to go idle — a deadlock until the exchange
# times out — so the test process loads modules eagerly.  (One process per GPU, the production layout, has no such
# coupling.)  Must be set before the CUDA driver is initialised.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "gat-recommendation_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
