#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dataset_featurize.py tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v11a.log 2>&1; echo "pytest new exit $?"
tail -15 gpurun_out/pytest_v11a.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v11.log 2>&1; echo "pytest full exit $?"
tail -4 gpurun_out/pytest_v11.log | cut -c1-250
timeout 300 python scripts/prof_dataprep.py > gpurun_out/dataprep_v11.txt 2>&1; echo "dataprep $?"; cat gpurun_out/dataprep_v11.txt | tail -8
timeout 600 python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; echo "bench exit $?"; tail -3 gpurun_out/bench_v11.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v11.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"])
r = d["roofline"]; print(r["kernel"]); print("frac", r["frac"], "achieved", r["achieved"], "parts", r["launch_ms_parts"], "fwd", r["conv_forward"])
print(d["kernel_ms_per_step"]); print(d["cpu_baseline"])
PY
