import csv
rows = list(csv.DictReader(l for l in open("gpurun_out/wgrad_ncu.csv") if l.startswith('"')))
out = {}
for r in rows:
    out.setdefault(r["ID"], {"k": r["Kernel Name"][:24], "g": r.get("Grid Size", "")})[r["Metric Name"][:16]] = r["Metric Value"]
seen = set()
for k, v in out.items():
    key = (v["k"], v["g"])
    if key in seen:
        continue
    seen.add(key)
    print(v)
