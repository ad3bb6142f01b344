#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_round_end.sh
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_v16_ref.json 2> gpurun_out/bench_v16_ref.err; echo "bench ref exit $?"; head -c 400 gpurun_out/bench_v16_ref.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 3 -c 1 -f -o gpurun_out/wgrad_v16 python scripts/probe_wgrad.py > gpurun_out/ncu_wgrad_v16.log 2>&1; echo "ncu wgrad exit $?"
