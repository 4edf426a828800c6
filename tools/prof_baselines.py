"""Runs a few GAT / GraphSAGE training steps on bench-shaped session batches — the command that
`ncu --metrics gpu__time_duration.sum` wraps to get their launch lists.
    python tools/prof_baselines.py [gat|sage] [sessions]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from etpgt_b200 import ops, optim, synth  # noqa: E402
from etpgt_b200.model import create_gat, create_graphsage  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "gat"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
data = synth.generate()
keys = synth.sorted_edge_keys(data)
dev = torch.device("cuda")
db = [h.to_device(dev) for h in bench.make_host_batches(data, keys, 0, batch, 2, seed=1, pin=False)]
torch.manual_seed(0)
model = (create_gat(82174, 256, 256, 3, 4, dropout=0.1) if kind == "gat"
         else create_graphsage(82174, 256, 256, 3, dropout=0.1)).to(dev)
opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
model.train()


def step(b):
    loss = ops.sampled_loss(model(b), model.item_embedding, b.target_item, b.negative_items, "bpr")[0]
    opt.zero_grad()
    loss.backward()
    opt.step()


for i in range(4):
    step(db[i % 2])
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    step(db[i % 2])
e.record()
torch.cuda.synchronize()
print(f"{kind} batch {batch}: {a.elapsed_time(e) / 10:.3f} ms/step")
