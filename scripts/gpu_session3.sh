#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/ea_timing.txt
for m in pyg bonds; do timeout 300 python scripts/prof_edgeattn.py $m bf16 >> gpurun_out/ea_timing.txt 2>&1; done
timeout 300 python scripts/prof_edgeattn.py bonds fp32 >> gpurun_out/ea_timing.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python scripts/prof_step.py pyg > gpurun_out/step_profile_pyg.txt 2>&1
tail -3 gpurun_out/smoke.log; tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/ea_timing.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; head -45 gpurun_out/step_profile_pyg.txt | cut -c1-200
