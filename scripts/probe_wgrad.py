"""Weight-gradient reductions of the trunk backward at config-2 sizes: current torch.mm forms vs alternatives."""
import torch
dev = "cuda"
n, na, H = 98304, 8544, 256
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3
bf = torch.bfloat16
xb = torch.randn(n, H, device=dev, dtype=bf)
tail = torch.randn(n, 2 * H, device=dev, dtype=bf)
dxr = tail[:, :H]                                   # strided view, as in the trunk
dbuf = torch.randn(na, 7 * H, device=dev, dtype=bf)
abar = torch.randn(na, 4 * H, device=dev, dtype=bf)
dagg = torch.randn(na, H, device=dev, dtype=bf)
out = torch.empty(H, H, device=dev)
print("dWs  mm(dxr.t(), xb) f32            us", timeit(lambda: torch.mm(dxr.t(), xb, out_dtype=torch.float32, out=out)))
print("dWs  mm(xb.t(), dxr).t() f32        us", timeit(lambda: torch.mm(xb.t(), dxr, out_dtype=torch.float32)))
dxc = dxr.contiguous()
print("dWs  contiguous dxr                 us", timeit(lambda: torch.mm(dxc.t(), xb, out_dtype=torch.float32)))
print("dWs  bf16 out                       us", timeit(lambda: torch.mm(dxc.t(), xb)))
x3, d3 = xb.view(8, n // 8, H), dxc.view(8, n // 8, H)
print("dWs  bmm 8 chunks + sum             us", timeit(lambda: torch.bmm(d3.transpose(1, 2), x3, out_dtype=torch.float32).sum(0)))
x3, d3 = xb.view(32, n // 32, H), dxc.view(32, n // 32, H)
print("dWs  bmm 32 chunks + sum            us", timeit(lambda: torch.bmm(d3.transpose(1, 2), x3, out_dtype=torch.float32).sum(0)))
o7 = torch.empty(7 * H, H, device=dev)
print("dW7  mm(dbuf.t(), xb[:na]) f32      us", timeit(lambda: torch.mm(dbuf.t(), xb[:na], out_dtype=torch.float32, out=o7)))
print("dW7  mm(xb[:na].t(), dbuf).t()      us", timeit(lambda: torch.mm(xb[:na].t(), dbuf, out_dtype=torch.float32)))
print("gfull mm(dagg.t(), abar) f32        us", timeit(lambda: torch.mm(dagg.t(), abar, out_dtype=torch.float32)))
print("gfull bmm per head                  us", timeit(lambda: torch.bmm(dagg.view(na, 4, 64).permute(1, 2, 0), abar.view(na, 4, H).transpose(0, 1), out_dtype=torch.float32)))
print("colsum-as-mv dxr                    us", timeit(lambda: torch.mm(torch.ones(1, n, device=dev, dtype=bf), dxc, out_dtype=torch.float32)))
dy = torch.randn(n, H, device=dev)
w = torch.randn(2 * H, H, device=dev, dtype=bf) * 0.05
print("dx   addmm(dy, tail, w) f32         us", timeit(lambda: torch.addmm(dy, tail, w, out_dtype=torch.float32)))
ws = torch.randn(H, H, device=dev, dtype=bf) * 0.05; bs = torch.randn(H, device=dev, dtype=bf)
print("xr   addmm(bs, xb, ws.t())          us", timeit(lambda: torch.addmm(bs, xb, ws.t())))
for ch in (16, 48, 64, 96, 128, 192):
    if n % ch: continue
    x3, d3 = xb.view(ch, n // ch, H), tail.view(ch, n // ch, 2 * H)[:, :, :H]
    print(f"dWs  bmm {ch} chunks (strided dxr) + sum us", timeit(lambda: torch.bmm(d3.transpose(1, 2), x3, out_dtype=torch.float32).sum(0)))
# both weight gradients of the tail at once: [dx_r | df]^T xb  (only the first H rows are needed, but one GEMM reads tail once)
print("dWs  mm(tail.t(), xb) f32 [2H x H]  us", timeit(lambda: torch.mm(tail.t(), xb, out_dtype=torch.float32)))
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_elasticity_predictor_b200 import ops
ob = torch.empty(H, device=dev); o7b = torch.empty(7 * H, device=dev)
for tc in (True, False):
    ops.WGRAD_TC = tc
    print(f"wgrad(dxr, xb)   + bias  tc={tc}  us", timeit(lambda: ops.wgrad(dxr, xb, out, ob)))
    print(f"wgrad(dbuf, xb[:na]) + bias tc={tc} us", timeit(lambda: ops.wgrad(dbuf, xb[:na], o7, o7b)))
