#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_parity.py > gpurun_out/parity_report.txt 2>&1
for m in pyg bonds; do timeout 300 python scripts/prof_conv.py $m bf16 >> gpurun_out/conv_timing.txt 2>&1; done
timeout 300 python scripts/prof_conv.py bonds fp32 >> gpurun_out/conv_timing.txt 2>&1
timeout 600 python scripts/prof_step.py pyg > gpurun_out/step_profile_pyg.txt 2>&1
# ncu: full capture of the conv kernels (plain run of the same command succeeded just above)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 12 -c 6 -o gpurun_out/conv_pyg -f python scripts/prof_conv.py pyg bf16 3 > gpurun_out/ncu_conv_pyg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 12 -c 6 -o gpurun_out/conv_bonds -f python scripts/prof_conv.py bonds bf16 3 > gpurun_out/ncu_conv_bonds.log 2>&1
cat gpurun_out/parity_report.txt gpurun_out/conv_timing.txt; head -60 gpurun_out/step_profile_pyg.txt; tail -3 gpurun_out/ncu_conv_pyg.log
