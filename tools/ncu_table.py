#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep as a small CSV (what profiles/*_kernels.csv hold).
  python tools/ncu_table.py gpurun_out/r2_bam_kernels.ncu-rep > profiles/r2_bam_kernels.csv"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
cols = [w for w in want if w in H]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + cols)
w.writerow(["unit"] + [U[H.index(c)] for c in cols])
for r in rows[2:]:
    w.writerow([r[H.index("Kernel Name")].split("(")[0]] + [r[H.index(c)] for c in cols])
