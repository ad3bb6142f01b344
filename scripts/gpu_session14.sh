#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_lgattn.py pyg 3 active > gpurun_out/plain_lg_v8.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lgattn_(fwd|bwd)_kernel" -s 6 -c 2 -f -o gpurun_out/lg_v8 python scripts/prof_lgattn.py pyg 3 active > gpurun_out/ncu_lg_v8.log 2>&1
echo "ncu full exit $?"; cat gpurun_out/plain_lg_v8.log | tail -8; tail -2 gpurun_out/ncu_lg_v8.log | cut -c1-200
