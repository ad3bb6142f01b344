"""tcgen05 + TMA node projection (csrc/proj_tc.cu) against cuBLAS (torch.addmm) at the config-2 shapes, L2 flushed.
usage: python scripts/prof_proj_tc.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_elasticity_predictor_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for m, n in ((98304, 256), (8544, 1792), (98304, 1792), (8192, 256), (8192, 1792)):
    x = (torch.randn(m, 256, device=dev, generator=g)).to(torch.bfloat16)
    w = (torch.randn(n, 256, device=dev, generator=g) * 0.06).to(torch.bfloat16)
    b = torch.randn(n, device=dev, generator=g).to(torch.bfloat16)
    res = {}
    for name, fn in (("tcgen05", lambda: ops.linear_lp(x, w, b)), ("cuBLAS", lambda: torch.addmm(b, x, w.t()))):
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        res[name] = statistics.median(ts)
    byts = 2 * (m * 256 + n * 256 + m * n)
    print(f"M={m:6d} N={n:5d}: tcgen05 {res['tcgen05']:7.1f} us ({byts / res['tcgen05'] / 1e3:6.0f} GB/s, {2 * m * n * 256 / res['tcgen05'] / 1e6:6.1f} TFLOP/s)"
          f"   cuBLAS {res['cuBLAS']:7.1f} us ({byts / res['cuBLAS'] / 1e3:6.0f} GB/s)")

# one conv block's projections: x_r over all rows + q|k|v|qt over the active prefix -- one tcgen05 launch vs two cuBLAS GEMMs
for n, na in ((98304, 8544), (98304, 98304), (8192, 8192)):
    x = (torch.randn(n, 256, device=dev, generator=g)).to(torch.bfloat16)
    w8 = (torch.randn(2048, 256, device=dev, generator=g) * 0.06).to(torch.bfloat16)
    b8 = torch.randn(2048, device=dev, generator=g).to(torch.bfloat16)
    res = {}
    for name, fn in (("tcgen05", lambda: ops.block_projections(x, w8, b8, na)),
                     ("cuBLAS", lambda: (torch.addmm(b8[1792:], x, w8[1792:].t()), torch.addmm(b8[:1792], x[:na], w8[:1792].t())))):
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        res[name] = statistics.median(ts)
    byts = 2 * (n * 256 + 2048 * 256 + n * 256 + na * 1792)
    print(f"block n={n:6d} active={na:6d}: tcgen05 one launch {res['tcgen05']:7.1f} us ({byts / res['tcgen05'] / 1e3:6.0f} GB/s)   "
          f"cuBLAS two GEMMs {res['cuBLAS']:7.1f} us")
