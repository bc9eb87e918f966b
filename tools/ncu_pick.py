"""Print a few headline metrics per kernel launch from an ncu --page raw --csv dump.   python tools/ncu_pick.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__inst_executed.sum", "launch__grid_size"]
for r in rows[2:]:
    print(r[ix["Kernel Name"]][:60])
    for k in keys:
        if k in ix:
            print("   %-95s %s %s" % (k, r[ix[k]], rows[1][ix[k]]))
