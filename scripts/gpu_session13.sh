#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edgeattn.py tests/test_gpu_engine.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_v8.log 2>&1; echo "pytest exit $?"
grep -v "^  \|^$" gpurun_out/pytest_v8.log | tail -8 | cut -c1-300
for m in pyg bonds; do timeout 300 python scripts/prof_lgattn.py $m 10; done > gpurun_out/lg_timing_v8.txt 2>&1; cat gpurun_out/lg_timing_v8.txt | grep "lgattn\|angle"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/b.err; echo "bench $?"
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['others'])"
