"""Time the line-graph conv core (fwd, bwd) alone at BASELINE config-2 size; used under ncu too.
usage: python scripts/prof_conv.py [pyg|bonds] [bf16|fp32] [iters]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops
from bench import conv_bytes

mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
dt = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
b = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode)
n, e, H, h = b.edge_index.size(1), b.lg_edge_index.size(1), 256, 4
plan = pkg.build_plan(b.lg_edge_index.to(dev), n)
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = (torch.randn(n, H, device=dev, generator=g).to(dt).requires_grad_(True) for _ in range(3))
ee = torch.randn(e, H, device=dev, generator=g).to(dt).requires_grad_(True)
gout = torch.randn(n, H, device=dev, generator=g)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ops.STATS.events = True
for it in range(iters + 3):
    if it == 3:
        torch.cuda.synchronize(); ops.STATS.reset()
    flush.zero_()
    out = pkg.conv_core(q, k, v, ee, plan, h)
    flush.zero_()
    out.backward(gout)
torch.cuda.synchronize()
d = ops.STATS.durations_ms()
s = 2 if dt == torch.bfloat16 else 4
for name, fwd in (("conv_fwd", True), ("conv_bwd", False)):
    ms = statistics.median(x[0] for x in d[name])
    by = conv_bytes(n, e, H, h, s, fwd)
    print(f"{mode} {dt} {name}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/6452.8:.3f} of measured peak)  bytes={by}")
