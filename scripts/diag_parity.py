"""Per-tensor parity report of the CUDA path vs the fp64 oracle (and the fp32 oracle's own noise)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model_ref
import gnn_elasticity_predictor_b200 as pkg

dev = "cuda"
torch.set_num_threads(os.cpu_count() or 8)
ref32 = model_ref.build_hetero(hidden=256, layers=4, heads=4, seed=42)
ref64 = copy.deepcopy(ref32).double()
ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 256, 4, 4, 0.0), 2).to(dev)
ours.load_state_dict(ref32.state_dict())

def run_ref(model, batch, dbl):
    model.zero_grad()
    b = copy.copy(batch)
    if dbl:
        for k in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot"):
            setattr(b, k, getattr(batch, k).double())
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    m, l = model(b)
    loss = model_ref.gaussian_nll_loss(m, l, tz.double() if dbl else tz)
    loss.backward()
    return m.detach(), l.detach(), loss.detach(), {k: p.grad.detach().double() for k, p in model.named_parameters() if p.grad is not None}

def run_ours(batch, autocast):
    ours.zero_grad()
    b = batch.to(dev)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        m, l = ours(b)
        loss = pkg.gaussian_nll_loss(m.float(), l.float(), pkg.zscore_targets(b.y, b.num_graphs))
    loss.backward()
    return m.detach().double().cpu(), l.detach().double().cpu(), loss.detach().double().cpu(), {k: p.grad.detach().double().cpu() for k, p in ours.named_parameters() if p.grad is not None}

for inc in ("pyg", "bonds"):
    batch = pkg.synthetic_batch(64, 16, 12, seed=0, lg_inc=inc)
    m64, l64, loss64, g64 = run_ref(ref64, batch, True)
    m32, l32, loss32, g32 = run_ref(ref32, batch, False)
    gmax = max(float(v.abs().max()) for v in g64.values())
    for tag, (m, l, loss, g) in (("cuda fp32", run_ours(batch, False)), ("cuda bf16", run_ours(batch, True)), ("cpu  fp32 oracle", (m32.double(), l32.double(), loss32.double(), g32))):
        fm = float((m - m64).abs().max() / m64.abs().max()); fl = float((l - l64).abs().max() / l64.abs().max())
        rows = []
        for k, w in g64.items():
            err = float((g[k] - w).abs().max()); mx = float(w.abs().max())
            cos = float(torch.nn.functional.cosine_similarity(g[k].flatten(), w.flatten(), dim=0))
            rows.append((err / max(mx, 1e-3 * gmax), err / gmax, cos, k))
        rows.sort(reverse=True)
        print(f"[{inc}] {tag}: mean rel {fm:.2e} logvar rel {fl:.2e} loss rel {abs(float(loss - loss64)) / abs(float(loss64)):.2e}  "
              f"worst grad rel(floor 1e-3 gmax) {rows[0][0]:.2e} ({rows[0][3]})  worst err/gmax {max(r[1] for r in rows):.2e}  min cos {min(r[2] for r in rows if 'lin_key.bias' not in r[3]):.5f}")
        for r in rows[:4]:
            print("      %.2e  err/gmax %.2e  cos %.5f  %s" % r)
