"""torch.profiler breakdown of the bench step (kernel-level table by CUDA time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import dp
from bench import ARCH

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
torch.manual_seed(0)
model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
model.base.compute_dtype = torch.bfloat16
model.train()
bucket = dp.FlatGradBucket(model.parameters())
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode).to(dev)
tz = pkg.zscore_targets(batch.y, 256)

def step():
    bucket.zero()
    model.base.build_plans(batch)
    mean, logvar = model(batch)
    loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz)
    loss.backward()
    dp.global_grad_clip(bucket, 5.0)
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
