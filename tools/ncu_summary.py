"""Summarise an `ncu -i rep --page raw --csv` export: the metrics the profiling recipe names."""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "sm__cycles_elapsed.max"]
# every tensor-pipe / TMEM / clock metric the capture holds (names differ between pipes: take them by pattern)
PATTERNS = ("sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__inst_executed_pipe_uniform", "tmem", "sm__cycles_active.avg",
            "gpc__cycles_elapsed.avg.per_second", "sm__throughput.avg", "gpu__dram_throughput", "lts__throughput")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("---")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-88s %s %s" % (w, r[i], units[i]))
    for i, name in enumerate(hdr):
        if name not in WANT and any(pt in name for pt in PATTERNS):
            print("  %-88s %s %s" % (name, r[i], units[i]))
