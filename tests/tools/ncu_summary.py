#!/usr/bin/env python
"""profiles/ summary of one captured kernel: the headline metrics of `ncu --page raw --csv` plus the per-line table of ncu_lines.py.
   python tests/tools/ncu_summary.py <raw.csv> <lines.txt> > profiles/NAME.txt"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print("  %-86s %s %s" % (k, d[k], u.get(k, "")))
print()
print(open(sys.argv[2]).read())
