#!/usr/bin/env python
"""Summarise an .ncu-rep into a small CSV (one row per captured launch) for profiles/.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_sectors_srcunit_tex_op_red.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [w for w in WANT if w in idx]
with open(sys.argv[2], "w", newline="") as f:
    wr = csv.writer(f)
    wr.writerow([f"{c} [{units[idx[c]]}]" if units[idx[c]] else c for c in cols])
    for r in rows[2:]:
        wr.writerow([r[idx[c]] for c in cols])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
