"""Probe torch/cuBLAS features and GEMM timings used by the block restructure (config-2 line-graph sizes)."""
import torch, time
dev = "cuda"
n, H = 98304, 256
def timeit(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
xb = torch.randn(n, H, device=dev, dtype=torch.bfloat16)
w4 = torch.randn(4 * H, H, device=dev, dtype=torch.bfloat16) * 0.05
w8 = torch.randn(8 * H, H, device=dev, dtype=torch.bfloat16) * 0.05
b4 = torch.randn(4 * H, device=dev, dtype=torch.bfloat16)
b8 = torch.randn(8 * H, device=dev, dtype=torch.bfloat16)
dy = torch.randn(n, H, device=dev)
d4 = torch.randn(n, 4 * H, device=dev, dtype=torch.bfloat16)
d8 = torch.randn(n, 8 * H, device=dev, dtype=torch.bfloat16)
print("proj4 addmm us", timeit(lambda: torch.addmm(b4, xb, w4.t())))
print("proj8 addmm us", timeit(lambda: torch.addmm(b8, xb, w8.t())))
print("dx4 mm us", timeit(lambda: d4 @ w4))
print("dx8 mm us", timeit(lambda: d8 @ w8))
try:
    r = torch.mm(d8, w8, out_dtype=torch.float32)
    print("mm out_dtype ok", r.dtype, "us", timeit(lambda: torch.mm(d8, w8, out_dtype=torch.float32)))
except Exception as e:
    print("mm out_dtype FAIL", e)
try:
    r = torch.addmm(dy, d8, w8, out_dtype=torch.float32)
    ref = dy + (d8.float() @ w8.float())
    print("addmm f32 self + out_dtype ok", r.dtype, float((r - ref).abs().max() / ref.abs().max()),
          "us", timeit(lambda: torch.addmm(dy, d8, w8, out_dtype=torch.float32)))
except Exception as e:
    print("addmm out_dtype FAIL", str(e)[:300])
print("dy + mm us", timeit(lambda: dy + d8 @ w8))
print("dW4 us", timeit(lambda: d4.t() @ xb))
print("dW8 us", timeit(lambda: d8.t() @ xb))
try:
    print("dW8 f32 out us", timeit(lambda: torch.mm(d8.t(), xb, out_dtype=torch.float32)))
except Exception as e:
    print("dW8 f32 FAIL", str(e)[:200])
print("colsum d8 torch us", timeit(lambda: d8.sum(0, dtype=torch.float32)))
# bmm shapes
h, C = 4, 64
wc3 = torch.randn(h, C, H, device=dev, dtype=torch.bfloat16) * 0.05
q3 = d4[:, :H].unflatten(1, (h, C)).transpose(0, 1)
abar = torch.randn(h, n, H, device=dev, dtype=torch.bfloat16)
print("qt bmm us", timeit(lambda: torch.bmm(q3, wc3)))
print("agge bmm us", timeit(lambda: torch.bmm(abar, wc3.transpose(1, 2))))
print("dwc bmm us", timeit(lambda: torch.bmm(q3.transpose(1, 2), abar)))
try:
    print("dwc bmm f32 us", timeit(lambda: torch.bmm(q3.transpose(1, 2), abar, out_dtype=torch.float32)))
except Exception as e:
    print("bmm f32 FAIL", str(e)[:200])
# strided-out baddbmm
dq3 = d4[:, :H].unflatten(1, (h, C)).transpose(0, 1)
before = d4.clone()
try:
    torch.baddbmm(dq3, abar, wc3.transpose(1, 2), out=dq3)
    want = before[:, :H].float() + torch.bmm(abar.float(), wc3.float().transpose(1, 2)).transpose(0, 1).reshape(n, H)
    print("baddbmm strided out err", float((d4[:, :H].float() - want).abs().max() / want.abs().max()),
          "others untouched", bool(torch.equal(d4[:, H:], before[:, H:])),
          "us", timeit(lambda: torch.baddbmm(dq3, abar, wc3.transpose(1, 2), out=dq3)))
except Exception as e:
    print("baddbmm strided FAIL", str(e)[:200])
# foreach copy into flat views
ps = [torch.randn(256, 256, device=dev) for _ in range(130)]
flat = torch.zeros(130 * 65536, device=dev)
views = [flat[i * 65536:(i + 1) * 65536].view(256, 256) for i in range(130)]
print("foreach_copy us", timeit(lambda: torch._foreach_copy_(views, ps)))
