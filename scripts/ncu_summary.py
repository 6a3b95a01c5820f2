#!/usr/bin/env python
"""Condense an .ncu-rep into the per-launch metrics the roofline discussion uses (profiles/*.md)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
print("| kernel | " + " | ".join(w.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", "") for w in want if w in idx) + " |")
print("|---|" + "---|" * len([w for w in want if w in idx]))
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    vals = []
    for w in want:
        if w in idx:
            v = r[idx[w]]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v + " " + units[idx[w]])
    print("| " + name + " | " + " | ".join(vals) + " |")
