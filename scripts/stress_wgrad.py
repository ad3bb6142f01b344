"""Randomised stress of the tcgen05 + TMA weight-gradient kernel: many (K, M, stride) shapes against fp64, each run twice."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_elasticity_predictor_b200 import ops
ops.WGRAD_TC, ops.WGRAD_TC_MAX_M = True, -1
rng = random.Random(0)
dev = "cuda"
worst = 0.0
for it in range(300):
    k = rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 4095, 4096, 8544, 20000, rng.randint(1, 30000)])
    m = 8 * rng.randint(1, 224)
    pad = rng.choice([0, 8, 256])
    a_full = (torch.randn(k, m + pad, device=dev) * 0.5).to(torch.bfloat16)
    a = a_full[:, :m]
    b_full = torch.randn(k, 256 + rng.choice([0, 8]), device=dev).to(torch.bfloat16)
    b = b_full[:, :256]
    w, s = torch.full((m, 256), float("nan"), device=dev), torch.full((m,), float("nan"), device=dev)
    ops.wgrad(a, b, w, s)
    w2, s2 = torch.empty_like(w), torch.empty_like(s)
    ops.wgrad(a, b, w2, s2)
    ww, ss = a.double().t() @ b.double(), a.double().sum(0)
    ew = float((w.double() - ww).abs().max()) / max(float(ww.abs().max()), 1e-30)
    es = float((s.double() - ss).abs().max()) / max(float(ss.abs().max()), 1e-30)
    worst = max(worst, ew, es)
    if not (ew < 5e-6 and es < 5e-6 and torch.equal(w, w2) and torch.equal(s, s2)):
        print("FAIL", it, k, m, pad, ew, es, flush=True)
        sys.exit(1)
torch.cuda.synchronize()
print("wgrad stress OK: 300 shapes, worst rel err", worst)
