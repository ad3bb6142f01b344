"""Soak: N training steps of one member at BASELINE config 2 through TrainStep (CUDA-graph replay, dropout on) fed from the
device-resident store; checks that the loss stays finite and goes down, and reports steps/s.  usage: soak_train.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import dataset, engine
from bench import ARCH

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = torch.Generator().manual_seed(0)
store = dataset.DeviceGraphStore([pkg.make_crystal(32, 12, gen) for _ in range(512)], dev)
model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
model.base.compute_dtype = torch.bfloat16
model.train()
ts = engine.TrainStep(model, lr=3e-4, weight_decay=1e-4)
losses = []
t0 = None
for i in range(steps):
    ids = torch.randperm(512, generator=gen)[:256]
    st = ts.static_inputs(probe) if i > 3 else None
    b = store.collate(ids.numpy(), out=None if st is None else st[0])
    if i == 0:
        probe = b
    tz = pkg.zscore_targets(b.y, b.num_graphs)
    loss, _, _ = ts.step(b, tz)
    if i % 100 == 0 or i == steps - 1:
        losses.append(float(loss))
        if i == 100:
            torch.cuda.synchronize(); t0 = time.perf_counter(); i0 = i
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("loss every 100 steps:", " ".join(f"{x:.3f}" for x in losses))
print(f"{steps} steps, replays {ts.replays}, eager {ts.eager_steps}; {(steps - 1 - i0) / dt:.1f} steps/s = {(steps - 1 - i0) * 256 / dt:.0f} graphs/s incl. host loop")
assert all(x == x and abs(x) < 1e4 for x in losses), "loss diverged / NaN"
assert min(losses[-5:]) < losses[0], "loss did not go down"
print("SOAK OK")
