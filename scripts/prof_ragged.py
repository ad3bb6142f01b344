"""Ragged batches (crystals of different size, as a real dataset gives): eager step vs shape-bucket padding + graph replay.
usage: python scripts/prof_ragged.py [n_graphs] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine, batching
from bench import ARCH

n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
batches = []
for i in range(6):
    sizes = torch.randint(24, 41, (n_graphs,), generator=gen).tolist()
    batches.append(pkg.collate([pkg.make_crystal(a, 12, gen) for a in sizes], lg_inc="pyg").to(dev))
print("batch sizes:", [b.sizes for b in batches[:3]], "buckets:", {tuple(batching.bucket_shape(b).values()) for b in batches})
tzs = [pkg.zscore_targets(b.y, b.num_graphs) for b in batches]
for name, kw in (("eager (ragged shapes)", dict(graph=False)), ("pad_to_buckets + CUDA-graph replay", dict(graph=True, pad_to_buckets=True))):
    torch.manual_seed(1)
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
    model.base.compute_dtype = torch.bfloat16
    model.train()
    ts = engine.TrainStep(model, **kw)
    for i in range(12):
        ts.step(batches[i % 6], tzs[i % 6])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        loss = ts.step(batches[i % 6], tzs[i % 6])[0]
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / steps
    print(f"{name:40s} {dt * 1e3:7.2f} ms/step  {n_graphs / dt:9.0f} graphs/s   graphs captured {len(ts._captured)} replays {ts.replays} eager {ts.eager_steps}  loss {float(loss):.4f}")
