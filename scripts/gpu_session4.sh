#!/bin/bash
# tests + bench + ncu launch list + ncu --set full of the streaming kernels (v2 baseline before the MMA rewrite)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 python scripts/prof_edgeattn.py pyg bf16 3 > gpurun_out/ea_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edgeattn -s 6 -c 3 -f -o gpurun_out/ea_v2 python scripts/prof_edgeattn.py pyg bf16 3 > gpurun_out/ncu_ea.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 1 --members 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; tail -5 gpurun_out/ncu_ea.log; tail -3 gpurun_out/ncu_launches.log
