"""torch.profiler (CUPTI) breakdown of the bench step: per-kernel GPU time of one eager TrainStep at config-2 size.
usage: python scripts/prof_step.py [pyg|bonds] [n_steps]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from torch.autograd import DeviceType
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine
from bench import ARCH

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
model.base.compute_dtype = torch.bfloat16
model.train()
ts = engine.TrainStep(model, graph=False)
batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode).to(dev)
tz = pkg.zscore_targets(batch.y, 256)
for _ in range(3):
    ts.step(batch, tz)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(n_steps):
        ts.step(batch, tz)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        t, c = agg.get(e.name, (0.0, 0))
        agg[e.name] = (t + e.device_time, c + 1)
total = sum(t for t, _ in agg.values())
print(f"# {mode}: {n_steps} eager steps, GPU kernel time {total / n_steps / 1e3:.3f} ms/step, "
      f"{sum(c for _, c in agg.values()) // n_steps} kernels/step")
for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / n_steps:10.1f} us {100 * t / total:5.1f}%  x{c // n_steps:4d}  {name[:150]}")
