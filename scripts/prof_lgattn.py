"""Time the line-graph kernels of the fused trunk alone at BASELINE config-2 size (also the ncu target).
usage: python scripts/prof_lgattn.py [pyg|bonds] [iters] [active]   ("active": run on the active prefix of bond rows only,
as the trunk program does with PyG-collated batches)"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops
from bench import lgattn_bytes, edgeattn_bytes, load_peaks

mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev, dt = "cuda", torch.bfloat16
b = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode)
n_all, e, H, h = b.edge_index.size(1), b.lg_edge_index.size(1), 256, 4
plan = pkg.build_plan(b.lg_edge_index.to(dev), n_all)
n = b.lg_active_rows if (len(sys.argv) > 3 and sys.argv[3] == "active") else n_all
g = torch.Generator(device=dev).manual_seed(0)
proj = (torch.randn(n, 8 * H, device=dev, generator=g) * 0.5).to(dt)
dbuf = torch.empty(n, 9 * H, device=dev, dtype=dt)
q, k, v, xr = (proj[:, i * H:(i + 1) * H] for i in range(4))
qt = proj[:, 4 * H:].unflatten(1, (h, H)).transpose(0, 1)
dq, dk, dv, dxr = (dbuf[:, i * H:(i + 1) * H] for i in range(4))
bbar = dbuf[:, 4 * H:8 * H].unflatten(1, (h, H)).transpose(0, 1)
gt = (torch.randn(h, n, H, device=dev, generator=g) * 0.1).to(dt)
dagg = torch.randn(n, H, device=dev, generator=g)
dlp = dagg.to(dt)
cvec = torch.randn(H, device=dev, generator=g) * 0.1
a = torch.rand(e, 11, device=dev, generator=g)
w1 = torch.randn(H, 11, device=dev, generator=g) * 0.5
b1 = torch.randn(H, device=dev, generator=g) * 0.2
a_csr = ops.pack_angles(a, plan)
abar_rows = torch.empty(n, h, H, device=dev, dtype=dt)
x32 = torch.randn(n, H, device=dev, generator=g)
wb, gm, bl = (torch.randn(s, device=dev, generator=g) * 0.1 for s in (3 * H, H, H))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ops.STATS.events = True
for it in range(iters + 3):
    if it == 3:
        torch.cuda.synchronize(); ops.STATS.reset()
    flush.zero_()
    aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, h, 0.15, 1, it, abar=abar_rows.transpose(0, 1))
    agge = torch.bmm(abar, torch.randn(h, 64, H, device=dev, dtype=dt).transpose(1, 2))
    flush.zero_()
    y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cvec, s, h, xr, x32, wb, gm, bl, 1e-5, 0.15, 2, it, True)
    flush.zero_()
    dagg2, dlp2, dpar = ops.raw_gate_ln_bwd3(dagg, agg, xr, wb, gm, bl, beta, mean, rstd, s, h, dxr, 0.15, 2, it,
                                             dy2=dbuf[:, 8 * H:])
    flush.zero_()
    coef = ops.raw_lgattn_bwd(dagg, dlp, agg, q, k, v, qt, gt, cvec, a_csr, w1, b1, m, z, plan, h, dq, dk, dv, bbar, 0.15, 1, it)
    flush.zero_()
    ops.colsum(dbuf[:, :8 * H])
    if it % 4 == 0:
        flush.zero_()
        ops.raw_lg_angle_grad(a_csr, w1, b1, plan, [coef] * 4, [qt] * 4, [gt] * 4)
torch.cuda.synchronize()
d = ops.STATS.durations_ms()
peak, _ = load_peaks()
by = {"lgattn_fwd": lgattn_bytes("lgattn_fwd", n, e, H, h, 2), "lgattn_bwd_dst": lgattn_bytes("lgattn_bwd_dst", n, e, H, h, 2),
      "edgeattn_bwd_src": edgeattn_bytes("edgeattn_bwd_src", n, e, H, h, 2) - 2 * H * n,
      "gate_ln_fwd": (4 + 2 + 2 + 4 + 4 + 4 + 2) * H * n, "gate_ln_bwd": (4 + 2 + 4 + 2 + 4 + 2 + 2) * H * n,
      "colsum": 2 * 8 * H * n, "lg_angle_grad": 4 * (2 * 2 * h * H * n + 32 * e) + 32 * e}
for name in ("lgattn_fwd", "lgattn_bwd_dst", "edgeattn_bwd_src", "gate_ln_fwd", "gate_ln_bwd", "colsum", "lg_angle_grad"):
    ms = statistics.median(x[0] for x in d[name])
    print(f"{mode} n={n} {name:18s} {ms * 1e3:8.1f} us   algorithmic {by[name] / 1e6:8.1f} MB  -> {by[name] / ms / 1e6:7.0f} GB/s "
          f"({by[name] / ms / 1e6 / peak:.3f} of measured HBM peak {peak:.0f})")
