"""Summarise `ncu -i X.ncu-rep --page raw --csv` output as markdown tables (one per captured launch).
usage: ncu_summary.py raw.csv [raw2.csv ...] > summary.md"""
import csv, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.min", "smsp__cycles_active.max",
           "smsp__cycles_active.avg", "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        print(f"### `{d.get('Kernel Name', '?')[:100]}`  (launch id {d.get('ID')}, {path.split('/')[-1]})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for m in METRICS:
            if m in d and d[m] != "":
                print(f"| `{m}` | {d[m]} | {u.get(m, '')} |")
        print()
