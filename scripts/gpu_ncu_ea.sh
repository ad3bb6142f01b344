#!/bin/bash
# usage: gpu_ncu_ea.sh <tag> [pyg|bonds]  -- ncu --set full of the streaming line-graph kernels at config-2 size
mkdir -p gpurun_out
MODE=${2:-bonds}
timeout 300 python scripts/prof_edgeattn.py $MODE bf16 3 > gpurun_out/ea_plain_$1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-edgeattn} -s ${KSKIP:-6} -c ${KCOUNT:-2} -f -o gpurun_out/ea_$1 python scripts/prof_edgeattn.py $MODE bf16 3 > gpurun_out/ncu_ea_$1.log 2>&1
cat gpurun_out/ea_plain_$1.log; tail -3 gpurun_out/ncu_ea_$1.log
