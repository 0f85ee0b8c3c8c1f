#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel launch, `ncu --set full`) and/or an ncu launch list CSV into
markdown for profiles/.  Usage:
    python tools/ncu_summary.py --rep gpurun_out/x.ncu-rep --launches gpurun_out/launches.csv > profiles/rNN_x.md
"""
import argparse
import csv
import subprocess
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_tex_op_atom.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
    "lts__t_sector_hit_rate.pct",
]


def rep_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--launches")
    ap.add_argument("--title", default="ncu summary")
    a = ap.parse_args()
    print(f"# {a.title}\n")
    for rep in a.rep:
        rows = rep_rows(rep)
        h, u = rows[0], rows[1]
        ki = h.index("Kernel Name")
        for v in rows[2:]:
            print(f"## `{v[ki]}`  ({rep})\n")
            print("| metric | value | unit |\n|---|---|---|")
            for i, name in enumerate(h):
                if name in KEYS or (name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio")):
                    print(f"| {name} | {v[i]} | {u[i]} |")
            print()
    if a.launches:
        rows = list(csv.reader(open(a.launches)))
        hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        h = rows[hdr]
        ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows[hdr + 1:]:
            if len(r) <= vi:
                continue
            v = float(r[vi].replace(",", ""))
            v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
            agg[r[ki][:90]][0] += 1
            agg[r[ki][:90]][1] += v
        tot = sum(v[1] for v in agg.values())
        print(f"## launch list ({a.launches}; cold-cache, serialised: compare shares, not absolutes)\n")
        print("| kernel | launches | total us | share |\n|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
