"""BASELINE config 5 shape: ensemble inference (5 members, forward only, eval mode) on batches of 256 x 32-atom cells.
The reference's predict.py runs fp32 without autocast (predict.py:604-623); the bf16 regime is what eval_epoch_hetero uses
(train.py:757).   usage: python scripts/prof_inference.py [pyg|bonds] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ensemble
from bench import ARCH

mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
members = []
for i in range(5):
    torch.manual_seed(42 + 1007 * i)
    members.append(pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev).eval())
batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode).to(dev)
for dt in (torch.bfloat16, torch.float32):
    for m in members:
        m.base.compute_dtype = dt
    for _ in range(3):
        batch._alignn_plans = None
        ensemble.ensemble_forward(members, batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters):
        batch._alignn_plans = None                      # plan build (sorts) is part of every batch
        mean_z, var_z, std_z = ensemble.ensemble_forward(members, batch)
    torch.cuda.synchronize(); dt_s = (time.perf_counter() - t0) / iters
    print(f"{mode} {str(dt):15s} 5-member ensemble forward: {dt_s * 1e3:7.2f} ms per batch of 256 -> {256 / dt_s:8.0f} structures/s "
          f"({256 * 5 / dt_s:8.0f} member-forwards/s)   mean|mean_z| {float(mean_z.abs().mean()):.4f}")
