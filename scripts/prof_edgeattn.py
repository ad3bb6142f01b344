"""Time the streaming line-graph kernels alone at BASELINE config-2 size (also the ncu target).
usage: python scripts/prof_edgeattn.py [pyg|bonds] [bf16|fp32] [iters]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops
from bench import edgeattn_bytes

mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
dt = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
b = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode)
n, e, H, h = b.edge_index.size(1), b.lg_edge_index.size(1), 256, 4
plan = pkg.build_plan(b.lg_edge_index.to(dev), n)
g = torch.Generator(device=dev).manual_seed(0)
proj = (torch.randn(n, 4 * H, device=dev, generator=g) * 0.5).to(dt)
dproj = torch.empty_like(proj)
q, k, v = (proj[:, i * H:(i + 1) * H] for i in range(3))
dq, dk, dv = (dproj[:, i * H:(i + 1) * H] for i in range(3))
qt = (torch.randn(h, n, H, device=dev, generator=g) * 0.1).to(dt)
gt = (torch.randn(h, n, H, device=dev, generator=g) * 0.1).to(dt)
feat = torch.relu(torch.randn(e, H, device=dev, generator=g)).to(dt)
dagg = torch.randn(n, H, device=dev, generator=g)
cvec = torch.randn(H, device=dev, generator=g) * 0.1
df = torch.empty_like(feat)
a = torch.rand(e, 11, device=dev, generator=g)
w1 = torch.randn(H, 11, device=dev, generator=g) * 0.5
b1 = torch.randn(H, device=dev, generator=g) * 0.2
a_csr = ops.pack_angles(a, plan) if dt == torch.bfloat16 else None
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ops.STATS.events = True
for it in range(iters + 3):
    if it == 3:
        torch.cuda.synchronize(); ops.STATS.reset()
    flush.zero_()
    aggv, abar, m, z, s = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, h, 0.0, 0, 0)
    if a_csr is not None:
        flush.zero_()
        ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, h, 0.0, 0, 0)
    flush.zero_()
    ops.raw_edgeattn_bwd(dagg, None, aggv, q, k, v, qt, gt, cvec, feat, m, z, plan, h, dq, dk, dv, df if it % 2 else None, df,
                         False, 0.0, 0, 0)
torch.cuda.synchronize()
d = ops.STATS.durations_ms()
sb = 2 if dt == torch.bfloat16 else 4
if "lgattn_fwd" in d:
    ms = statistics.median(x[0] for x in d["lgattn_fwd"])
    by = edgeattn_bytes("edgeattn_fwd", n, e, H, h, sb)
    print(f"{mode} {dt} lgattn_fwd (h1 recomputed in-kernel): {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s of stored-feature "
          f"algorithmic bytes ({by/ms/1e6/6452.8:.3f} of measured peak)  bytes={by}")
for name in ("edgeattn_fwd", "edgeattn_bwd_dst", "edgeattn_bwd_src"):
    for accum in ((False, True) if name == "edgeattn_bwd_dst" else (False,)):
        xs = [x[0] for x in d[name] if name != "edgeattn_bwd_dst" or bool(x[1][5]) == accum]
        ms = statistics.median(xs)
        by = edgeattn_bytes(name, n, e, H, h, sb, accum)
        print(f"{mode} {dt} {name}{' (accumulate)' if accum else ''}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s "
              f"({by/ms/1e6/6452.8:.3f} of measured peak)  bytes={by}")
