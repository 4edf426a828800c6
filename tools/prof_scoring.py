"""Runs only the tensor-core scoring kernel (evaluation shape) a few times — the command that
`ncu --set full -k regex:score_topk_tc` wraps (see profiles/)."""

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))

from etpgt_b200 import ops  # noqa: E402

sessions, items, dim, k = int(sys.argv[1]) if len(sys.argv) > 1 else 23861, 82174, 256, 20
g = torch.Generator().manual_seed(0)
sess = ops.to_bf16((torch.randn(sessions, dim, generator=g) * 0.1).cuda())
table = ops.to_bf16((torch.randn(items, dim, generator=g) * 0.1).cuda())
for _ in range(3):
    ops.score_topk(sess, table, k, precision="bf16")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.score_topk(sess, table, k, precision="bf16")
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"score_topk bf16 {sessions}x{items}x{dim} k={k}: {ms:.3f} ms, {2.0 * sessions * items * dim / ms / 1e9:.1f} TFLOP/s")
