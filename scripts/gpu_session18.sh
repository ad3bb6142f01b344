#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_dataset_featurize.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v12a.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_v12a.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo "bench exit $?"; tail -3 gpurun_out/bench_v12.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v12.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "store", d["e2e_device_store"])
PY
# ncu launch list of the same bench command (eager so that every kernel is its own launch), after the plain run above exited 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v12.csv python bench.py --no-graph --members 1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_v12.log 2>&1; echo "ncu launches exit $?"
python scripts/launch_summary.py gpurun_out/launches_v12.csv 8 > gpurun_out/launches_v12_step.txt 2>&1; head -30 gpurun_out/launches_v12_step.txt
