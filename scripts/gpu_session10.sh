#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench exit $?"
timeout 600 python bench.py --lg-inc bonds --no-cpu-baseline > gpurun_out/bench_v6_bonds.json 2> gpurun_out/bench_v6_bonds.err; echo "bench bonds exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v6_ref.json 2> gpurun_out/bench_v6_ref.err; echo "bench ref exit $?"
grep -v "^  \|^$" gpurun_out/pytest_gpu_full.log | tail -15 | cut -c1-300; tail -3 gpurun_out/smoke.log
for f in v6 v6_bonds v6_ref; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$f.json"))
    print("$f", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "roof", (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("kernel"))
except Exception as e:
    print("$f", "ERR", e); print(open("gpurun_out/bench_$f.err").read()[-1500:])
PY
done
