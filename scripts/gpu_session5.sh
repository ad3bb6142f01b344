#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edgeattn.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_ea.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ea.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
rm -f gpurun_out/ea_timing.txt
for m in pyg bonds; do timeout 300 python scripts/prof_edgeattn.py $m bf16 >> gpurun_out/ea_timing.txt 2>&1; done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -30 gpurun_out/pytest_ea.log; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/ea_timing.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
