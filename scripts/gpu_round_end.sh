#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v13.log 2>&1; echo "pytest full exit $?"; tail -6 gpurun_out/pytest_v13.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; echo "bench exit $?"; tail -3 gpurun_out/bench_v13.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v13.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "store", d["e2e_device_store"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["conv_forward"]["frac"])
print(d["kernel_ms_per_step"])
PY
