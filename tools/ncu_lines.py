#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Line No")
iL, iI, iT, iS = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
agg, src, cur = {}, {}, None
for r in rows:
    if not r or r is hdr or len(r) < len(hdr):
        continue
    if r[iL] != "":
        try:
            cur = int(r[iL])
        except ValueError:
            continue
        src[cur] = r[1]
        continue
    try:
        a = agg.setdefault(cur, [0, 0, 0])
        a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iS])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
print("total warp instructions", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:5d} inst {v[0]:9d} {100 * v[0] / tot:5.1f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / tots:5.1f}%  {src[k][:120]}")
