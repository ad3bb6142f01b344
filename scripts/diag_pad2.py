import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops
from test_gpu_edgeattn import lg_case, HEADS, H, C, DEV
n, e = 257, 6000
index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1 = lg_case(n, e, 31, 300)
pad = 700
index_p = torch.cat([index, torch.full((2, pad), -1, dtype=index.dtype)], dim=1)
a_p = torch.cat([a, torch.zeros(pad, a.size(1), device=a.device)])
res = {}
for name, idx, aa in (("plain", index, a), ("padded", index_p, a_p)):
    plan = pkg.build_plan(idx.to(DEV), n)
    a_csr = ops.pack_angles(aa, plan)
    aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, 0.0, 0, 3)
    agge = torch.bmm(abar.float(), wc.to(DEV).transpose(1, 2)).transpose(0, 1).reshape(n, H)
    agg = aggv + agge + (cvec.view(HEADS, C) * s.unsqueeze(-1)).reshape(n, H)
    dproj = torch.zeros(n, 4 * H, dtype=torch.bfloat16, device=DEV)
    dq, dk, dv = (dproj[:, t * H:(t + 1) * H] for t in range(3))
    bbar = torch.empty(HEADS, n, H, dtype=torch.bfloat16, device=DEV)
    coef = ops.raw_lgattn_bwd(dagg, dagg.to(torch.bfloat16), agg, q, k, v, qt, gt, cvec, a_csr, w1, b1, m, z, plan, HEADS, dq, dk, dv, bbar, 0.0, 0, 3)
    torch.cuda.synchronize()
    res[name] = dict(aggv=aggv, abar=abar, dq=dq.clone(), dk=dk.clone(), dv=dv.clone(), bbar=bbar, coef=coef[:e].clone(),
                     rowptr_t=plan.rowptr_t.clone(), col_t=plan.col_t[:e].clone(), eid_t=plan.eid_t[:e].clone(), pos=ops.csc_positions(plan)[:e].clone(),
                     rowptr=plan.rowptr.clone(), eid=plan.eid[:e].clone())
for kk in res["plain"]:
    x, y = res["plain"][kk], res["padded"][kk]
    print(kk, "equal" if torch.equal(x, y) else f"DIFF max {float((x.float() - y.float()).abs().max()):.4g}")
