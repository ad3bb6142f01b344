#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_cfg.log 2>&1; echo "pytest exit $?"
grep -v "^  \|^$" gpurun_out/pytest_cfg.log | tail -12 | cut -c1-300
timeout 900 python bench.py --workload config4 --steps 5 --warmup 3 --no-cpu-baseline --members 1 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit $?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_cfg4.json")); print({k: d[k] for k in ("value", "ms_per_step")}, d["config"]["workload"], (d.get("e2e") or {}).get("value"), d["roofline"]["frac"] if d.get("roofline") else None)
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_cfg4.err").read()[-2000:])
PY
nvidia-smi --query-gpu=memory.used --format=csv
