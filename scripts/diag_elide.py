import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, copy
import gnn_elasticity_predictor_b200 as pkg
from oracle import model_ref
DEV = "cuda"
CTOR = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=2, heads=4)
torch.manual_seed(8)
ref = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(dropout=0.0, **CTOR), 2)
m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **CTOR), 2).to(DEV)
m.load_state_dict(ref.state_dict())
m.train(); m.base.compute_dtype = torch.bfloat16
host = pkg.synthetic_batch(24, 16, 12, seed=7, lg_inc="pyg")
tz = pkg.zscore_targets(host.y, host.num_graphs)
ref64 = copy.deepcopy(ref).double()
b64 = copy.copy(host)
for k, v in host.tensors().items():
    if v.is_floating_point(): setattr(b64, k, v.double())
rm, rl = ref64(b64)
model_ref.gaussian_nll_loss(rm, rl, tz.double()).backward()
want = {k: p.grad for k, p in ref64.named_parameters() if p.grad is not None}
batch = host.to(DEV)
out = {}
for name, elide, fused in (("elide", True, True), ("full", False, True), ("perblock", False, False)):
    m.base.elide_isolated = elide; m.base.fused_trunk = fused
    m.zero_grad(set_to_none=True)
    mean, logvar = m(batch)
    pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz.to(DEV)).backward()
    out[name] = {k: p.grad.double().cpu() for k, p in m.named_parameters() if p.grad is not None}
gmax = max(float(v.abs().max()) for v in want.values())
print("tensor | |g|max/gmax | err(elide) err(full) err(perblock) vs fp64 | elide-full")
for k, w in want.items():
    sc = max(float(w.abs().max()), 1e-3 * gmax)
    e = [float((out[n][k] - w).abs().max()) / sc for n in ("elide", "full", "perblock")]
    d = float((out["elide"][k] - out["full"][k]).abs().max()) / sc
    if max(e) > 5e-3 or d > 2e-3:
        print(f"{k:45s} {float(w.abs().max())/gmax:8.1e}  {e[0]:.4f} {e[1]:.4f} {e[2]:.4f}   {d:.4f}")
