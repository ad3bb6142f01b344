#!/bin/bash
# Round-2 ncu captures of the SHIPPED kernels (each command first exits 0 without ncu).  Outputs: gpurun_out/r02_*.ncu-rep/.csv
mkdir -p gpurun_out
K='lgattn_(fwd|bwd)_kernel|conv_bwd_src_kernel|lg_angle_grad_kernel|gate_ln_(fwd|bwd)_kernel'
python scripts/prof_lgattn.py pyg 3 active > gpurun_out/r02_plain_lg_pyg.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 14 -c 7 -f -o gpurun_out/r02_lg_pyg python scripts/prof_lgattn.py pyg 3 active > gpurun_out/r02_ncu_lg_pyg.log 2>&1; echo "ncu pyg exit $?"
python scripts/prof_lgattn.py bonds 3 > gpurun_out/r02_plain_lg_bonds.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:"lgattn_(fwd|bwd)_kernel|conv_bwd_src_kernel|lg_angle_grad_kernel" -s 8 -c 4 -f -o gpurun_out/r02_lg_bonds python scripts/prof_lgattn.py bonds 3 > gpurun_out/r02_ncu_lg_bonds.log 2>&1; echo "ncu bonds exit $?"
python scripts/prof_proj_tc.py > gpurun_out/r02_plain_proj.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"proj_tc_kernel" -s 100 -c 3 -f -o gpurun_out/r02_proj_tc python scripts/prof_proj_tc.py > gpurun_out/r02_ncu_proj.log 2>&1; echo "ncu proj exit $?"
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bonds --no-graph > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bonds --no-graph > gpurun_out/r02_bench_ncu.log 2>&1; echo "ncu launches exit $?"
ls -la gpurun_out/r02_*.ncu-rep gpurun_out/r02_launches_bench.csv | tail -6
