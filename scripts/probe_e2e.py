"""Where does the end-to-end loop lose time in graph mode?  Variants of the e2e loop, wall-clock per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine
from bench import ARCH

dev = torch.device("cuda", 0)
torch.manual_seed(0)
M = 2
steppers = []
for m in range(M):
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
    model.base.compute_dtype = torch.bfloat16
    model.train()
    steppers.append(engine.TrainStep(model, graph=True))
host = [pkg.synthetic_batch(256, 32, 12, seed=i).pin_memory() for i in range(2)]
devb = host[0].to(dev)
tz = pkg.zscore_targets(devb.y, 256)
for i in range(3 * M + 2):
    steppers[i % M].step(devb, tz)
torch.cuda.synchronize()
copy_stream = torch.cuda.Stream()
main = torch.cuda.current_stream()

def run(name, n, upload, sync_every, same_stream=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cpu_replay = 0.0
    nxt = None
    for i in range(n):
        if upload:
            if same_stream:
                b = host[i % 2].to(dev, non_blocking=True)
            else:
                if nxt is None:
                    with torch.cuda.stream(copy_stream):
                        nxt = (host[0].to(dev, non_blocking=True), torch.cuda.Event()); nxt[1].record(copy_stream)
                b, ev = nxt
                main.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    nxt = (host[(i + 1) % 2].to(dev, non_blocking=True), torch.cuda.Event()); nxt[1].record(copy_stream)
                for t in b.tensors().values():
                    t.record_stream(main)
        else:
            b = devb
        c0 = time.perf_counter()
        loss, mean, logvar = steppers[i % M].step(b, tz)
        cpu_replay += time.perf_counter() - c0
        if sync_every:
            float(loss)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:55s} {dt:7.3f} ms/step   cpu in step(): {cpu_replay / n * 1e3:6.3f} ms", flush=True)

n = 20
run("graph, resident batch, no per-step sync", n, False, False)
run("graph, resident batch, sync every step", n, False, True)
run("graph, upload on copy stream, no per-step sync", n, True, False)
run("graph, upload on copy stream, sync every step", n, True, True)
run("graph, upload on main stream, no per-step sync", n, True, False, same_stream=True)
# raw H2D time
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    b = host[i % 2].to(dev, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D alone: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per batch ({host[0].nbytes() / 1e6:.1f} MB)")
for s in steppers:
    s.use_graph = False
run("eager, resident batch, no per-step sync", n, False, False)
run("eager, upload on copy stream, no per-step sync", n, True, False)
run("eager, upload on copy stream, sync every step", n, True, True)
