"""Runs only the tensor-core scoring kernel (evaluation shape) a few times — the command that
`ncu --set full -k regex:score_dump_tc` wraps (see profiles/).
    python tools/prof_scoring.py [sessions] [normal|xavier|popular] [flush]"""

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))

from etpgt_b200 import ops  # noqa: E402

sessions = int(sys.argv[1]) if len(sys.argv) > 1 else 23861
mode = sys.argv[2] if len(sys.argv) > 2 else "normal"
flush_l2 = len(sys.argv) > 3
items, dim, k = 82174, 256, 20
g = torch.Generator().manual_seed(0)
sess = (torch.randn(sessions, dim, generator=g) * 0.1).cuda()
if mode == "normal":
    table = (torch.randn(items, dim, generator=g) * 0.1).cuda()
elif mode == "xavier":      # nn.init.xavier_uniform_ on [items, dim], row 0 = padding
    bound = (6.0 / (items + dim)) ** 0.5
    table = ((torch.rand(items, dim, generator=g) * 2 - 1) * bound).cuda()
    table[0] = 0
else:                       # trained-like: a popularity direction shared by many items
    table = (torch.randn(items, dim, generator=g) * 0.05).cuda()
    table += torch.randn(1, dim, generator=g).cuda() * torch.rand(items, 1, generator=g).cuda()
sess_h, table_h = ops.to_bf16(sess), ops.to_bf16(table)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.score_topk(sess_h, table_h, k, precision="bf16")
torch.cuda.synchronize()
total = 0.0
for _ in range(5):
    if flush_l2:
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.score_topk(sess_h, table_h, k, precision="bf16")
    b.record()
    torch.cuda.synchronize()
    total += a.elapsed_time(b)
ms = total / 5
print(f"score_topk bf16 {sessions}x{items}x{dim} k={k} [{mode}{' flush' if flush_l2 else ''}]: {ms:.3f} ms, "
      f"{2.0 * sessions * items * dim / ms / 1e9:.1f} TFLOP/s")
