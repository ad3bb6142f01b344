#!/bin/bash
# ncu --set full captures of the line-graph kernels (pyg active prefix, and bonds), the tcgen05 forward and the collate kernels
mkdir -p gpurun_out
timeout 300 python scripts/prof_lgattn.py pyg 3 active > gpurun_out/plain_lg_v14.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_lg_v14.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lgattn_(fwd|bwd)_kernel|conv_bwd_src_kernel|lg_angle_grad_kernel|gate_ln_(fwd|bwd)_kernel" -s 14 -c 7 -f -o gpurun_out/lg_v14_pyg python scripts/prof_lgattn.py pyg 3 active > gpurun_out/ncu_lg_v14_pyg.log 2>&1; echo "ncu pyg exit $?"
timeout 1200 ncu --set full --clock-control none -k regex:"lgattn_(fwd|bwd)_kernel|conv_bwd_src_kernel|lg_angle_grad_kernel" -s 8 -c 4 -f -o gpurun_out/lg_v14_bonds python scripts/prof_lgattn.py bonds 3 > gpurun_out/ncu_lg_v14_bonds.log 2>&1; echo "ncu bonds exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lgattn_fwd_tc_kernel" -s 8 -c 1 -f -o gpurun_out/tc_v14 python scripts/check_lgattn_tc.py full > gpurun_out/ncu_tc_v14.log 2>&1; echo "ncu tc exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"collate_rows_kernel|collate_index_kernel|linegraph_fill_kernel|bond_features_kernel" -s 30 -c 12 -f -o gpurun_out/dataprep_v14 python scripts/prof_dataprep.py > gpurun_out/ncu_dataprep_v14.log 2>&1; echo "ncu dataprep exit $?"
ls -la gpurun_out/*.ncu-rep | tail -5
