"""Parity + timing of the tcgen05 line-graph attention forward (csrc/lgattn_tc.cu) against the mma.sync kernel
(csrc/lgattn.cu) on the same inputs.  usage: python scripts/check_lgattn_tc.py [small|full]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops
from test_gpu_edgeattn import lg_case, HEADS

DEV = "cuda"
mode = sys.argv[1] if len(sys.argv) > 1 else "small"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def run(tc, *args, **kw):
    ops.LGATTN_TC = tc
    out = ops.raw_lgattn_fwd(*args, **kw)
    torch.cuda.synchronize()
    return out


def compare(tag, n, e, seed, hub, p=0.0):
    index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1 = lg_case(n, e, seed, hub)
    plan = pkg.build_plan(index.to(DEV), n)
    a_csr = ops.pack_angles(a, plan)
    ref = run(False, q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
    print(f"[{tag}] n={n} e={e} hub={hub} p={p}: launching tc kernel", flush=True)
    got = run(True, q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
    errs = {name: rel(x, y) for name, x, y in zip(("aggv", "abar", "m", "z", "s"), got, ref)}
    print(f"[{tag}] rel err vs mma.sync kernel:", {k_: f"{v_:.2e}" for k_, v_ in errs.items()}, flush=True)
    ok = all(v_ < (2e-2 if k_ != "m" else 1e-5) for k_, v_ in errs.items())
    if not ok:
        for name, x, y in zip(("aggv", "abar", "m", "z", "s"), got, ref):
            d = (x.double() - y.double()).abs()
            flat = int(d.argmax())
            print("   ", name, "worst at", flat, "got", float(x.flatten()[flat]), "want", float(y.flatten()[flat]),
                  "nan", int(torch.isnan(x).sum()), flush=True)
    return ok


ok = True
ok &= compare("tiny", 9, 40, 3, 4)
ok &= compare("multi-row", 301, 7000, 21, 500)
ok &= compare("dropout", 301, 7000, 21, 500, p=0.2)
ok &= compare("long-rows", 64, 9000, 5, 3000)
print("PARITY", "OK" if ok else "FAILED", flush=True)

if mode == "full":
    for lg_inc in ("pyg", "bonds"):
        b = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=lg_inc).to(DEV)
        n = b.edge_index.size(1)
        plan = pkg.build_plan(b.lg_edge_index, n)
        na = b.lg_active_rows if lg_inc == "pyg" else n
        g = torch.Generator().manual_seed(0)
        mk = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(DEV).to(torch.bfloat16)
        proj = mk(na, 7 * 256)
        q, k, v = (proj[:, i * 256:(i + 1) * 256] for i in range(3))
        qt = proj[:, 3 * 256:].unflatten(1, (4, 256)).transpose(0, 1)
        w1 = (torch.randn(256, 11, generator=g) * 0.5).to(DEV); b1 = (torch.randn(256, generator=g) * 0.2).to(DEV)
        a_csr = ops.pack_angles(b.lg_edge_attr, plan)
        plan_a = plan          # q has `na` rows: the kernels visit rows [0, na) only (rows beyond are isolated)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
        res = {}
        for tc in (False, True):
            outs = run(tc, q, k, v, qt, a_csr, w1, b1, plan_a, HEADS, 0.15, 7, 3)
            ts = []
            for _ in range(10):
                flush.zero_()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(); ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan_a, HEADS, 0.15, 7, 3); t1.record()
                torch.cuda.synchronize(); ts.append(t0.elapsed_time(t1))
            ts.sort(); res[tc] = (ts[len(ts) // 2], outs)
        import ctypes
        from gnn_elasticity_predictor_b200 import _lib
        dbg = torch.zeros(16, dtype=torch.int64, device=DEV)
        lib = _lib.load()
        lib.alignn_lgattn_tc_debug.argtypes = [ctypes.c_void_p]
        lib.alignn_lgattn_tc_debug(ctypes.c_void_p(dbg.data_ptr()))
        run(True, q, k, v, qt, a_csr, w1, b1, plan_a, HEADS, 0.15, 7, 3)
        lib.alignn_lgattn_tc_debug(None)
        d = dbg.cpu().tolist()
        names = ["take", "P1 mma1+wait", "P2 cvt", "P3 wait K+sync", "MMA2+wait", "P4 softmax", "P5 wait V+sync", "MMA3+wait", "epilogue"]
        print("  CTA0 cycles per tile over", d[9], "tiles:", {nm: round(c / max(d[9], 1)) for nm, c in zip(names, d[:9])}, flush=True)
        print("    inside: prefetch issue (A, K gather, B2)", round(d[10] / d[9]), "| S load + select", round(d[11] / d[9]),
              "| max: match/redux/sync", round(d[12] / d[9]), "| (P4 value = the rest: exp, sums, P store) | V gather issue",
              round(d[13] / d[9]), flush=True)
        errs = {name: rel(x, y) for name, x, y in zip(("aggv", "abar", "m", "z", "s"), res[True][1], res[False][1])}
        print(f"config2 {lg_inc}: rows {na}, angles {plan.n_edges}: mma.sync {res[False][0] * 1e3:.1f} us, tcgen05 {res[True][0] * 1e3:.1f} us; "
              f"rel err {({k_: f'{v_:.1e}' for k_, v_ in errs.items()})}", flush=True)
