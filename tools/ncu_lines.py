#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples, shared wavefronts) from an ncu report captured with
--import-source on:   python tools/ncu_lines.py report.ncu-rep [top=40]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname = "?"
agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0, ""])
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        col = {k: i for i, k in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (fname, int(r[0]))
    a = agg[key]
    def num(name):
        v = r[col[name]]
        try:
            return int(v.replace(",", ""))
        except ValueError:
            return 0
    if r[col["Address"]]:
        a[0] += num("Instructions Executed")
        a[1] += num("# Samples")
        a[2] += num("L1 Wavefronts Shared")
        a[3] += num("stall_long_sb")
        a[4] += num("stall_short_sb")
        a[5] += num("stall_barrier")
        a[6] += num("stall_wait")
    else:
        a[7] = r[1][:110]
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
print(f"{'file:line':22s} {'inst%':>6s} {'smp%':>6s} {'long':>6s} {'short':>6s} {'barr':>6s} {'wait':>6s} {'shwf/inst':>9s}  source")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f + ':' + str(l):22s} {100 * a[0] / ti:6.2f} {100 * a[1] / ts:6.2f} {100 * a[3] / ts:6.2f} {100 * a[4] / ts:6.2f} {100 * a[5] / ts:6.2f} {100 * a[6] / ts:6.2f} {a[2] / max(a[0], 1):9.2f}  {a[7]}")
