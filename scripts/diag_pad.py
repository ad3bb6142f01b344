import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import batching, fused
from test_batching import ragged_batch, CTOR
torch.manual_seed(3)
m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **CTOR), 2).cuda()
m.base.compute_dtype = torch.bfloat16
m.train()
host = ragged_batch(5, "bonds")
padded, mask = batching.pad_batch(host)
print(host.sizes, padded.sizes, host.lg_active_rows, padded.lg_active_rows)
nb = host.num_graphs
grads = {}
orig = fused.mlp2
def hooked(x, *a):
    y = orig(x, *a)
    y.register_hook(lambda g, n=x.size(1): grads.setdefault(cur[0], {}).__setitem__(n, g.detach().clone()))
    return y
fused.mlp2 = hooked
cur = [None]
res = {}
for name, b, mk in (("plain", host, None), ("padded", padded, mask)):
    cur[0] = name
    b = b.to("cuda")
    t = pkg.zscore_targets(b.y, b.num_graphs)
    m.zero_grad(set_to_none=True)
    mean, logvar = m(b)
    pkg.gaussian_nll_loss(mean.float(), logvar.float(), t, mask=None if mk is None else mk.cuda()).backward()
    res[name] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
m.base.compute_dtype = torch.float32
cur[0] = "fp32"
b = host.to("cuda"); t = pkg.zscore_targets(b.y, b.num_graphs)
m.zero_grad(set_to_none=True)
mean, logvar = m(b)
pkg.gaussian_nll_loss(mean.float(), logvar.float(), t).backward()
res["fp32"] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
E, N = host.sizes["E"], host.sizes["N"]
g32 = grads["fp32"][36].float()
print("edge grad: |plain-fp32|", float((grads["plain"][36].float() - g32).abs().max()), "|padded-fp32|", float((grads["padded"][36].float()[:E] - g32).abs().max()), "scale", float(g32.abs().max()))
gmax = max(float(v.abs().max()) for v in res["fp32"].values())
print("tensor: err(plain) err(padded) vs fp32, relative to max(|g|, 1e-2 gmax)")
for k, w in res["fp32"].items():
    sc = max(float(w.abs().max()), 1e-2 * gmax)
    ea, eb = float((res["plain"][k] - w).abs().max()) / sc, float((res["padded"][k] - w).abs().max()) / sc
    if max(ea, eb) > 0.02: print(f"{k:45s} {ea:.3f} {eb:.3f}")

for dim, n_real, lab in ((36, E, "edge"), (206, N, "node")):
    gp, gq = grads["plain"][dim].float(), grads["padded"][dim].float()
    print(lab, "real rows diff", float((gp - gq[:n_real]).abs().max()), "scale", float(gp.abs().max()), "pad rows max", float(gq[n_real:].abs().max()))
for k in res["plain"]:
    a, b = res["padded"][k], res["plain"][k]
    d = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)
    if d > 0.02: print(f"{k:45s} rel {d:.3f}")
