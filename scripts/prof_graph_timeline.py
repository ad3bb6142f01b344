"""Timeline of ONE replayed step graph at config-2 size (torch.profiler / CUPTI): every kernel with start, duration and
stream, the union-busy time of the step, idle gaps, and the per-stream sums -- what the critical path of the step is made of.
usage: python scripts/prof_graph_timeline.py [pyg|bonds] [out.txt]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine
from bench import ARCH

dev = torch.device("cuda", 0)
mode = sys.argv[1] if len(sys.argv) > 1 else "pyg"
out_path = sys.argv[2] if len(sys.argv) > 2 else None
torch.manual_seed(0)
model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **ARCH), 2).to(dev)
model.base.compute_dtype = torch.bfloat16
model.train()
ts = engine.TrainStep(model, graph=True)
batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=mode).to(dev)
tz = pkg.zscore_targets(batch.y, 256)
for _ in range(6):
    ts.step(batch, tz)
torch.cuda.synchronize()
assert ts.replays > 0
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ts.step(batch, tz)
    torch.cuda.synchronize()
    ts.step(batch, tz)
    torch.cuda.synchronize()
import json, tempfile
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
with open(path) as fh:
    trace = json.load(fh)
cats = collections.Counter(e.get("cat") for e in trace["traceEvents"])
print("# event categories:", dict(cats))
evs = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
evs.sort(key=lambda e: e["ts"])
# the second step = the second half of the activities (both steps launch the same work)
step = evs[len(evs) // 2:]
others = [e for e in trace["traceEvents"] if e.get("cat") not in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e
          and e["ts"] >= step[0]["ts"]]
t0 = step[0]["ts"]
t_end = max(e["ts"] + e["dur"] for e in step)
lines = [f"# {mode}: one replayed step graph: {len(step)} GPU activities, wall {t_end - t0:.1f} us"]
# union busy
busy, cur_s, cur_e = 0.0, None, None
for e in step:
    s, en = e["ts"], e["ts"] + e["dur"]
    if cur_s is None:
        cur_s, cur_e = s, en
    elif s <= cur_e:
        cur_e = max(cur_e, en)
    else:
        busy += cur_e - cur_s
        cur_s, cur_e = s, en
busy += cur_e - cur_s
per_stream = collections.defaultdict(float)
for e in step:
    per_stream[e["args"].get("stream")] += e["dur"]
lines.append(f"# union busy {busy:.1f} us, idle {t_end - t0 - busy:.1f} us; per stream kernel time: "
             + ", ".join(f"{k}: {v:.1f}" for k, v in sorted(per_stream.items(), key=lambda kv: -kv[1])))
agg = collections.OrderedDict()
for e in step:
    t, c = agg.get(e["name"], (0.0, 0))
    agg[e["name"]] = (t + e["dur"], c + 1)
lines.append("# ---- by kernel")
for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    lines.append(f"{t:10.1f} us  x{c:4d}  {name[:140]}")
lines.append("# ---- timeline (start us, dur us, gap-to-previous-end-on-any-stream, stream, name)")
prev_end = t0
for e in step:
    lines.append(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {e['ts'] - prev_end:7.1f}  s{e['args'].get('stream')}  {e['name'][:110]}")
    prev_end = max(prev_end, e["ts"] + e["dur"])
lines.append("# ---- non-kernel events during the step (cat, start us, dur us, name)")
for e in sorted(others, key=lambda e: e["ts"])[:200]:
    lines.append(f"{e.get('cat'):>18s} {e['ts'] - t0:9.1f} {e['dur']:9.1f}  {e['name'][:100]}")
text = "\n".join(lines)
if out_path:
    with open(out_path, "w") as fh:
        fh.write(text + "\n")
print("\n".join(lines[:70]))
