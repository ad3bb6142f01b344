"""torch.profiler breakdown of ONE fp32 eval forward (predict.py / evaluate.py regime) at config-2 size."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from torch.autograd import DeviceType
import gnn_elasticity_predictor_b200 as pkg
from bench import ARCH
dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
torch.manual_seed(0)
m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev).eval()
m.base.compute_dtype = torch.float32
batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode).to(dev)
with torch.no_grad():
    for _ in range(3):
        batch._alignn_plans = None
        m(batch)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        batch._alignn_plans = None
        m(batch)
        torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        t, c = agg.get(e.name, (0.0, 0)); agg[e.name] = (t + e.device_time, c + 1)
total = sum(t for t, _ in agg.values())
print(f"# fp32 eval forward, {mode}: GPU kernel time {total / 1e3:.3f} ms, {sum(c for _, c in agg.values())} kernels")
for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{t:10.1f} us {100 * t / total:5.1f}%  x{c:4d}  {name[:130]}")
