"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md cites.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {sys.argv[1]}: ncu --set full --clock-control none (per-launch values; cold-cache, serialised)")
for r in rows[2:]:
    print(f"\n## {r[idx['Kernel Name']].strip()}   (launch id {r[idx['ID']]})")
    for w in WANT:
        if w in idx:
            print(f"{w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
    try:
        rd, wr = float(r[idx['dram__bytes_read.sum']]), float(r[idx['dram__bytes_write.sum']])
        ru, wu = units[idx['dram__bytes_read.sum']], units[idx['dram__bytes_write.sum']]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        # units are per-column in the csv header row; values may be scaled per row by ncu -> print raw only
    except Exception:
        pass
