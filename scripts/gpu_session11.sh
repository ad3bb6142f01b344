#!/bin/bash
mkdir -p gpurun_out
for m in pyg bonds; do timeout 300 python scripts/prof_lgattn.py $m 10; done > gpurun_out/lg_timing_v6.txt 2>&1; echo "prof exit $?"
cat gpurun_out/lg_timing_v6.txt | tail -16
BENCH="python bench.py --no-graph --members 1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain_launch.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_v6.csv $BENCH > gpurun_out/ncu_launches_v6.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches_v6.log | cut -c1-300
timeout 300 python scripts/prof_lgattn.py pyg 2 > gpurun_out/plain_lg.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lgattn_(fwd|bwd)_kernel" -s 6 -c 2 -f -o gpurun_out/lg_v6 python scripts/prof_lgattn.py pyg 2 > gpurun_out/ncu_lg_v6.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_lg_v6.log | cut -c1-300
