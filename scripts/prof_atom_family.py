"""Atom-graph edge attention at config-2 size (8 192 atoms, 98 304 bonds): tensor-core (mma.sync) vs CUDA-core kernel family.
usage: python scripts/prof_atom_family.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops

dev, dt = "cuda", torch.bfloat16
for atoms, k_nb, graphs in ((32, 12, 256), (200, 16, 256)):
    b = pkg.synthetic_batch(graphs, atoms, k_nb, seed=0)
    n, e, H, h = b.x.size(0), b.edge_index.size(1), 256, 4
    plan = pkg.build_plan(b.edge_index.to(dev), n)
    g = torch.Generator(device=dev).manual_seed(0)
    proj = (torch.randn(n, 4 * H, device=dev, generator=g) * 0.5).to(dt)
    dproj = torch.empty_like(proj)
    q, k, v = (proj[:, i * H:(i + 1) * H] for i in range(3))
    dq, dk, dv = (dproj[:, i * H:(i + 1) * H] for i in range(3))
    qt = (torch.randn(h, n, H, device=dev, generator=g) * 0.1).to(dt)
    gt = (torch.randn(h, n, H, device=dev, generator=g) * 0.1).to(dt)
    feat = torch.randn(e, H, device=dev, generator=g).to(dt)
    dagg = torch.randn(n, H, device=dev, generator=g)
    cvec = torch.randn(H, device=dev, generator=g) * 0.1
    df = torch.empty_like(feat)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for mma in (True, False):
        ops.USE_MMA = mma
        ops.STATS.events = True
        for it in range(13):
            if it == 3:
                torch.cuda.synchronize(); ops.STATS.reset()
            flush.zero_()
            aggv, abar, m, z, s = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, h, 0.15, 1, it)
            flush.zero_()
            ops.raw_edgeattn_bwd(dagg, None, aggv, q, k, v, qt, gt, cvec, feat, m, z, plan, h, dq, dk, dv, None, df, False, 0.15, 1, it)
        torch.cuda.synchronize()
        d = ops.STATS.durations_ms()
        print(f"atoms/graph {atoms} nbrs {k_nb}: n={n} e={e}  {'mma.sync ' if mma else 'CUDA-core'}: "
              + "  ".join(f"{nm} {statistics.median(x[0] for x in d[nm]) * 1e3:7.1f} us" for nm in ("edgeattn_fwd", "edgeattn_bwd_dst", "edgeattn_bwd_src")), flush=True)
