#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.

  python profiles/summarize.py gpurun_out/r1_launches.csv gpurun_out/r1_all_kernels.ncu-rep r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]

# ---- launch list: device time of every launch of the bench command ------------------------------------
rows = list(csv.reader(open(launch_csv)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, D = rows[h], rows[h + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
per = collections.OrderedDict()
for r in D:
    if len(r) > vi:
        per.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")) / 1000.0)
with open(os.path.join(HERE, f"{tag}_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 3 --warmup 3\n")
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes.  Times in microseconds.\n")
    f.write("# Launches with the larger time belong to the 1M-record resident steps, the smaller ones to the e2e sub-batches.\n")
    f.write(f"{'kernel':44s} {'n':>4s} {'min':>9s} {'median':>9s} {'max':>9s}\n")
    big = {}
    for k, v in per.items():
        s = sorted(v)
        f.write(f"{k:44s} {len(v):4d} {s[0]:9.1f} {s[len(s) // 2]:9.1f} {s[-1]:9.1f}\n")
        if "exlr::" in k:
            big[k] = s[-1]
    tot = sum(big.values())
    f.write("\n# share of one 1M-record step (largest launch of each of our kernels)\n")
    for k, v in big.items():
        f.write(f"{k:44s} {v:9.1f} us  {100 * v / tot:5.1f} %\n")

# ---- full capture: key metrics per kernel --------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
cols = [w for w in want if w in H]
with open(os.path.join(HERE, f"{tag}_kernels.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + cols)
    w.writerow(["unit"] + [U[H.index(c)] for c in cols])
    traffic = {}
    for r in rows[2:]:
        name = r[H.index("Kernel Name")].split("(")[0]
        w.writerow([name] + [r[H.index(c)] for c in cols])
        def num(c):
            v, u = float(r[H.index(c)].replace(",", "")), U[H.index(c)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        traffic[name] = {"dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
                         "duration_us_under_ncu": float(r[H.index("gpu__time_duration.sum")].replace(",", ""))}
# DRAM bytes per launch of every kernel: bench.py reads the roofline kernel's entry into roofline.traffic
out = {}
for name, t in traffic.items():
    out[name.replace("void ", "").split("<")[0]] = {"source": f"profiles/{tag}_kernels.csv (ncu --set full, one launch on the 1M-record batch)",
                                                     "dram_bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"], **t}
path = os.path.join(HERE, "traffic.json")
try:
    old = json.load(open(path))
except Exception:
    old = {}
old.update(out)
old = {k: v for k, v in old.items()}
json.dump(old, open(path, "w"), indent=1)   # note: the k1_flat entry of a screened run is kernel 1c (listed mode), not the scan of everything
print(open(os.path.join(HERE, f"{tag}_launches_summary.txt")).read())
print(open(os.path.join(HERE, f"{tag}_kernels.csv")).read()[:3000])
