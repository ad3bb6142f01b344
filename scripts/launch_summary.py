"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel for the LAST 1/n-th of the launches
(= one step when the profiled command ran n identical steps).   usage: launch_summary.py launches.csv n_steps"""
import collections, csv, re, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
per = len(rows) // n_steps
last = rows[-per:]
agg = collections.OrderedDict()
for r in last:
    name = re.sub(r'<.*', '', r['Kernel Name'].replace('void ', ''))[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value']) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches captured, {n_steps} steps -> last step = {per} launches, {tot:.1f} us "
      "(ncu per-launch times: cold-cache, serialised -- compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% x{v[0]:4d}  {k}")
