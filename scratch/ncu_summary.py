"""ncu -i <rep> --page raw --csv  ->  one row per profiled launch with the columns quoted in profiles/README.md."""
import csv, subprocess, sys
COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
out = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    head, units = rows[0], rows[1]
    idx = [head.index(c) if c in head else None for c in COLS]
    if first:
        out.writerow(["Kernel Name"] + COLS)
        out.writerow([""] + [units[i] if i is not None else "" for i in idx])
        first = False
    name = head.index("Kernel Name")
    for r in rows[2:]:
        out.writerow([r[name]] + [r[i] if i is not None else "" for i in idx])
