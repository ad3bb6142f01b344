"""Times the gate / LayerNorm epilogue kernels alone at the line-graph shape of BASELINE config 2 (98 304 bond rows, 8 544
with an aggregate, H = 256, bf16, dropout 0.15), against the bytes they must move.  Inputs exceed L2 (> 300 MB per launch).

    python scripts/prof_gate_ln.py [eager] [n_rows] [active_rows]      (eager: two plain launches per case, for ncu)
"""
import sys

import torch

sys.path.insert(0, ".")
from gnn_elasticity_predictor_b200 import ops  # noqa: E402


EAGER = "eager" in sys.argv[1:]
if EAGER:
    sys.argv.remove("eager")


def timed(fn, reps=10, iters=10):
    """Kernel time only: `reps` calls captured into one CUDA graph (no host launch cost), replayed `iters` times."""
    fn()
    torch.cuda.synchronize()
    if EAGER:                           # under ncu: plain launches only
        fn()
        torch.cuda.synchronize()
        return float("nan")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (iters * reps) * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 98304
    na = int(sys.argv[2]) if len(sys.argv) > 2 else 8544
    hid, h = 256, 4
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s, dt=torch.float32: torch.randn(*s, device=dev, generator=g).to(dt)  # noqa: E731
    bf = torch.bfloat16
    aggv, agge, cv, s = r(na, hid), r(h, na, hid // h, dt=bf), r(hid), r(na, h).abs()
    xr, x32, xb = r(n, hid, dt=bf), r(n, hid), r(n, hid, dt=bf)
    wb, gm, bl = r(3 * hid) * 0.05, 1 + 0.1 * r(hid), 0.1 * r(hid)
    rs = ops.RNG_STEP
    for name, x, x_lp in (("fwd (fp32 residual in)", x32, None), ("fwd (bf16 residual in)", None, xb)):
        f = lambda: ops.raw_gate_ln_fwd2(aggv, agge, cv, s, h, xr, x, wb, gm, bl, 1e-5, 0.15, 7, 0, True, rs,  # noqa: E731
                                         agg_rows=na, x_lp=x_lp)
        us = timed(f)
        rd = n * hid * (2 + (4 if x is not None else 2)) + na * hid * (4 + 2) + na * h * 4
        wr = n * hid * (4 + 2) + n * 12 + na * hid * 4
        print(f"gate_ln_{name:24s} {us:7.1f} us   {(rd + wr) / 1e6:6.1f} MB  {(rd + wr) / us / 1e3:6.0f} GB/s")
    y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cv, s, h, xr, x32, wb, gm, bl, 1e-5, 0.15, 7, 0,
                                                          True, rs, agg_rows=na)
    dy = r(n, hid)
    tail = r(n, 2 * hid, dt=bf)
    for name, dy_, dy2 in (("bwd (dy + dy2)", dy, tail[:, hid:]), ("bwd (dy only)", dy, None), ("bwd (top: dy2 only)", None, tail[:, hid:])):
        f = lambda: ops.raw_gate_ln_bwd3(dy_, agg, xr, wb, gm, bl, beta, mean, rstd, s, h, tail[:, :hid], 0.15, 7, 0,  # noqa: E731
                                         rs, dy2=dy2, agg_rows=na)
        us = timed(f)
        rd = n * hid * (2 + (4 if dy_ is not None else 0) + (2 if dy2 is not None else 0)) + n * 12 + na * hid * 4 + na * h * 4
        wr = n * hid * 2 + na * hid * (4 + 2)
        print(f"gate_ln_{name:24s} {us:7.1f} us   {(rd + wr) / 1e6:6.1f} MB  {(rd + wr) / us / 1e3:6.0f} GB/s")


if __name__ == "__main__":
    main()
