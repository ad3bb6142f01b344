#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_model.py tests/test_gpu_edgeattn.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_v5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_v5.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_v5_eager.json 2> gpurun_out/bench_v5_eager.err; echo "bench eager exit $?"
grep -v "^  \|^$" gpurun_out/pytest_v5.log | tail -40 | cut -c1-400
cat gpurun_out/bench_v5.json; tail -5 gpurun_out/bench_v5.err; cat gpurun_out/bench_v5_eager.json; tail -5 gpurun_out/bench_v5_eager.err
