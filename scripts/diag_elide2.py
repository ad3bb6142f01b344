import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops, trunk
DEV = "cuda"
H, h = 256, 4
host = pkg.synthetic_batch(24, 16, 12, seed=7, lg_inc="pyg")
b = host.to(DEV)
n = b.edge_index.size(1)
plan = pkg.build_plan(b.lg_edge_index, n)
g = torch.Generator(device=DEV).manual_seed(0)
x32 = torch.randn(n, H, device=DEV, generator=g)
xb = x32.to(torch.bfloat16)
w8c = (torch.randn(8 * H, H, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
b8c = (torch.randn(8 * H, device=DEV, generator=g) * 0.1).to(torch.bfloat16)
wc3 = (torch.randn(h, 64, H, device=DEV, generator=g) * 0.05).to(torch.bfloat16)
cv = torch.randn(H, device=DEV, generator=g) * 0.1
wb, gm, bl = (torch.randn(s, device=DEV, generator=g) * 0.1 for s in (3 * H, H, H))
w1 = torch.randn(H, 11, device=DEV, generator=g) * 0.5
b1 = torch.randn(H, device=DEV, generator=g) * 0.2
a_csr = ops.pack_angles(b.lg_edge_attr, plan)
outs = {}
for name, act in (("full", -1), ("elide", host.lg_active_rows)):
    cfg = trunk.TrunkCfg(heads=h, n_layers=1, eps=[1e-5, 1e-5], p_attn=[0.0, 0.0], p_out=[0.0, 0.0], keys=[(0, 0, 0, 0)] * 2,
                         lg_plan=plan, g_plan=None, a_csr=a_csr, w1=w1, b1=b1, lg_active=act)
    y, ylp, st = trunk._block_forward(0, True, x32, xb, None, w8c, b8c, wc3, cv, wb, gm, bl, cfg, True, None)
    torch.cuda.synchronize()
    outs[name] = (y, st)
na = host.lg_active_rows
yf, sf = outs["full"]; ye, se = outs["elide"]
print("na", na, "n", n)
print("y diff rows<na", float((yf[:na] - ye[:na]).abs().max()), "rows>=na", float((yf[na:] - ye[na:]).abs().max()), "scale", float(yf.abs().max()))
for k in ("agg", "m", "z", "s", "beta", "mean", "rstd"):
    a, c = sf[k], se[k]
    m_ = min(a.size(0), c.size(0), na)
    print(k, "diff(<na)", float((a[:m_].float() - c[:m_].float()).abs().max()), "| tail full max", float(a[na:].float().abs().max()) if a.size(0) > na else None)
pf, pe = sf["proj"], se["proj"]
print("proj diff rows<na", float((pf[:na].float() - pe[:na].float()).abs().max()), "xr tail diff", float((pf[na:, 7 * H:].float() - pe[na:, 7 * H:].float()).abs().max()))
print("abar diff", float((sf["abar_rows"][:na].float() - se["abar_rows"].float()).abs().max()))
