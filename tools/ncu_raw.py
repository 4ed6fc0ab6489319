#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (ncu --set full): python tools/ncu_raw.py rep.ncu-rep"""
import csv
import subprocess
import sys

W = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
     'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
     'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
     'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
     'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
     'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum'] + \
    ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % x for x in
     ('long_scoreboard', 'short_scoreboard', 'barrier', 'wait', 'mio_throttle', 'math_pipe_throttle', 'not_selected', 'lg_throttle',
      'branch_resolving', 'no_instruction', 'dispatch_stall', 'drain', 'imc_miss', 'membar', 'sleeping', 'tex_throttle', 'selected')]
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
kn = H.index('Kernel Name')
for r in rows[2:]:
    print("==", r[kn][:60])
    for w in W:
        if w in H:
            print(f"   {w:88s} {r[H.index(w)]:>16s} {rows[1][H.index(w)]}")
