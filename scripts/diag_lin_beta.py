"""Where does the bf16 path's error on ``conv.lin_beta.weight`` gradients come from?  Config 1, seed 2: ours (bf16 / fp32),
the reference's own bf16 autocast (oracle on CPU) and the fp64 oracle, per lin_beta tensor and per 256-wide segment
(agg | x_r | agg - x_r): max-norm and L2 relative errors."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model_ref
import gnn_elasticity_predictor_b200 as pkg

dev = "cuda"
torch.set_num_threads(os.cpu_count() or 8)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ref = model_ref.build_hetero(hidden=256, layers=4, heads=4, seed=42)
ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 256, 4, 4, 0.0), 2).to(dev)
ours.load_state_dict(ref.state_dict(), strict=True)
batch = pkg.synthetic_batch(64, 16, 12, seed=seed, lg_inc="pyg")
tz = pkg.zscore_targets(batch.y, 64)

ref64 = copy.deepcopy(ref).double()
b64 = copy.copy(batch)
for k in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot"):
    setattr(b64, k, getattr(batch, k).double())
m, l = ref64(b64)
model_ref.gaussian_nll_loss(m, l, tz.double()).backward()
want = {k: p.grad.double() for k, p in ref64.named_parameters() if p.grad is not None}

with torch.autocast("cpu", dtype=torch.bfloat16):
    m, l = ref(batch)
    loss = model_ref.gaussian_nll_loss(m.float(), l.float(), tz)
loss.backward()
amp = {k: p.grad.double() for k, p in ref.named_parameters() if p.grad is not None}


def run_ours(autocast):
    ours.zero_grad()
    b = batch.to(dev)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        m, l = ours(b)
        loss = pkg.gaussian_nll_loss(m.float(), l.float(), pkg.zscore_targets(b.y, b.num_graphs))
    loss.backward()
    return {k: p.grad.detach().double().cpu() for k, p in ours.named_parameters() if p.grad is not None}


g16, g32 = run_ours(True), run_ours(False)
for k in sorted(want):
    if not k.endswith("lin_beta.weight"):
        continue
    w = want[k].flatten()
    print(k, "max|want| %.3e" % float(w.abs().max()))
    for tag, g in (("ours bf16", g16), ("ours fp32", g32), ("ref  AMP ", amp)):
        e = g[k].flatten() - w
        segs = " ".join("seg%d max %.2e l2 %.2e |w|max %.2e" % (s, float(e[256 * s:256 * s + 256].abs().max() / w.abs().max()),
                                                               float(e[256 * s:256 * s + 256].norm() / w[256 * s:256 * s + 256].norm()),
                                                               float(w[256 * s:256 * s + 256].abs().max())) for s in range(3))
        print("   %s: max %.3e  l2 %.3e   %s" % (tag, float(e.abs().max() / w.abs().max()), float(e.norm() / w.norm()), segs))
# L2 view of every tensor: ours bf16 vs reference AMP
print("\nL2 relative error per tensor (ours bf16 | reference AMP), worst 25 by ratio")
rows = []
for k, w in want.items():
    if k.endswith("lin_key.bias"):
        continue
    a, b = float((g16[k] - w).norm() / w.norm()), float((amp[k] - w).norm() / w.norm())
    rows.append((a / max(b, 1e-30), a, b, k))
rows.sort(reverse=True)
for r in rows[:25]:
    print("  ratio %.2f  ours %.3e  ref-amp %.3e  %s" % r)
