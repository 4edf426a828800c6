"""Host-side profile of the training step (cProfile over 20 steps at a small batch, where the step is
host-launch-bound)."""
import cProfile
import pstats
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "gat-recommendation_b200"))
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from etpgt_b200 import ops, optim, synth  # noqa: E402
from etpgt_b200.model import create_graph_transformer_optimized  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = synth.generate()
keys = synth.sorted_edge_keys(data)
dev = torch.device("cuda")
hb = bench.make_host_batches(data, keys, 0, batch, 2, seed=1, pin=True)
db = [h.to_device(dev) for h in hb]
model = create_graph_transformer_optimized(82174, 256, 256, 2, 2, dropout=0.1).to(dev)
model.laplacian_pe._cached_pe = bench.cached_pe(82174).to(dev)
opt = optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
model.train()


def step(b):
    sess = model(b)
    loss = ops.sampled_loss(sess, model.item_embedding, b.target_item, b.negative_items, "bpr")[0]
    opt.zero_grad()
    loss.backward()
    opt.step()


for i in range(6):
    step(db[i % 2])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(20):
    step(db[i % 2])
torch.cuda.synchronize()
print(f"batch {batch}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/step wall")
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    step(db[i % 2])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(40)
