#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v10_ref.json 2> gpurun_out/bench_v10_ref.err; echo "bench ref exit $?"
timeout 300 python scripts/prof_step.py pyg 3 > gpurun_out/step_profile_v10.txt 2> gpurun_out/step_profile_v10.err; echo "prof exit $?"
grep -v "^  \|^$" gpurun_out/pytest_gpu_full.log | tail -8 | cut -c1-300; tail -3 gpurun_out/smoke.log
head -c 600 gpurun_out/bench_v10.json; echo; cat gpurun_out/bench_v10_ref.json | head -c 1000; echo
head -60 gpurun_out/step_profile_v10.txt | cut -c1-160
