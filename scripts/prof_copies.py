"""Which aten ops issue device-to-device memcpys / copy kernels in one eager training step (input shapes + call sites)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine
from bench import ARCH
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
model.base.compute_dtype = torch.bfloat16
model.train()
ts = engine.TrainStep(model, graph=False)
batch = pkg.synthetic_batch(256, 32, 12, seed=0).to(dev)
tz = pkg.zscore_targets(batch.y, 256)
for _ in range(3):
    ts.step(batch, tz)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    ts.step(batch, tz)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if e.key in ("aten::copy_", "aten::clone", "aten::contiguous", "aten::_to_copy", "aten::cat", "aten::stack") and e.device_time_total > 15:
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:90], [s for s in e.stack if "gnn_elasticity" in s or "bench" in s][:2]))
for r in sorted(rows, reverse=True)[:24]:
    print(f"{r[0]:8.1f} us x{r[1]:3d} {r[2]:16s} {r[3]}  {r[4]}")
