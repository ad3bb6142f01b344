import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gnn_elasticity_predictor_b200 as pkg
from test_gpu_engine import _model
DEV = "cuda"
for trial in range(3):
    m = _model(seed=8, layers=2)
    m.base.compute_dtype = torch.bfloat16
    host = pkg.synthetic_batch(24, 16, 12, seed=7, lg_inc="pyg")
    batch = host.to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    out = {}
    for tag, elide in (("T", True), ("F", False), ("T2", True)):
        m.base.elide_isolated = elide
        m.zero_grad(set_to_none=True)
        mean, logvar = m(batch)
        pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz).backward()
        torch.cuda.synchronize()
        out[tag] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    gmax = max(float(g.abs().max()) for g in out["F"].values())
    for a, b in (("T", "F"), ("T", "T2")):
        worst = sorted(((float((out[a][k] - out[b][k]).abs().max()) / max(float(out[b][k].abs().max()), 1e-2 * gmax), k) for k in out[b]), reverse=True)[:4]
        print(trial, a, "vs", b, [(f"{r:.2e}", k) for r, k in worst])
