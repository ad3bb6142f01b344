#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edgeattn.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_ea.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ea.log
rm -f gpurun_out/ea_timing.txt
for m in pyg bonds; do timeout 300 python scripts/prof_edgeattn.py $m bf16 >> gpurun_out/ea_timing.txt 2>&1; done
grep -v "^  \|^$" gpurun_out/pytest_ea.log | tail -30 | cut -c1-300; cat gpurun_out/ea_timing.txt
