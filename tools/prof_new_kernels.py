"""Launches the kernels added in round 2 once each on representative sizes, for `ncu --set full`:
hub-row kernels of the fused TransformerConv on the zipf(1.5) graph (82,174 nodes, 1.43M directed edges, rows of
14,000 edges), the fused reduce-scatter + AdamW + all-gather of the item table (two ranks inside this process: the
"remote" half of the traffic stays on this GPU, so the NVLink part is NOT represented), the rank-major top-k merge.

    ncu --set full --clock-control none -k regex:"hub|dp_adam_table|merge_parts" -c 40 -o out python tools/prof_new_kernels.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from etpgt_b200 import _lib, ops, parallel  # noqa: E402
from etpgt_b200._lib import call, ptr, size, stream, workspace  # noqa: E402

dev = torch.device("cuda")
synth = bench.load_synth()
n = 82_174
zi, zj = synth.zipf_graph(n, 737_716)
src = torch.from_numpy(np.concatenate([zi, zj])).to(dev)
dst = torch.from_numpy(np.concatenate([zj, zi])).to(dev)
index = ops.GraphIndex(torch.stack([src, dst]), n)
qkvs = torch.randn(n, 1024, device=dev) * 0.3
w_beta = torch.randn(768, device=dev) * 0.1
index.hub_plan()
for _ in range(2):
    bench.time_tconv(qkvs, w_beta, index, dev, reps=1)

# fused table update, two ranks in this process, the 82,174 x 256 table
rows, dim = 82_174, 256
pad = (4 * rows * dim + 255) // 256 * 256
ctrl = int(_lib.size("etpgt_comm_control_bytes"))
comms = parallel.PeerComm.local_group(2, ctrl + 2 * pad)
m = [torch.zeros(rows, dim, device=dev) for _ in range(2)]
v = [torch.zeros(rows, dim, device=dev) for _ in range(2)]
for r, c in enumerate(comms):
    c.tensor(ctrl, (rows, dim)).normal_()
    c.tensor(ctrl + pad, (rows, dim)).normal_()
torch.cuda.synchronize()
for r, c in enumerate(comms):
    lo, hi = parallel.item_shard(rows, r, 2)
    call("etpgt_dp_adam_table", c.handle, ctrl + pad, ctrl, ptr(m[r]), ptr(v[r]), rows, dim, lo, hi, 1e-3, 0.9, 0.999, 1e-8,
         1e-5, 1, 1, stream())
torch.cuda.synchronize()

# rank-major merge of 8 shards' candidates, 23,861 sessions, k = 20
total, k, parts = 23_861, 20, 8
val_bytes = (total * k * 4 + 255) // 256 * 256
block = val_bytes + total * k * 8
buf = torch.empty(parts * block, dtype=torch.uint8, device=dev)
for p in range(parts):
    buf[p * block:p * block + total * k * 4].view(torch.float32).copy_(torch.randn(total * k, device=dev).sort(descending=True)[0])
    buf[p * block + val_bytes:(p + 1) * block].view(torch.int64).copy_(torch.randint(0, 82174, (total * k,), device=dev))
targets = torch.randint(0, 82174, (total,), device=dev)
ops.topk_merge_parts(buf, parts, block, val_bytes, total, k, 0, total, targets)
torch.cuda.synchronize()
print("done")
