#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics we track per kernel launch."""
import csv, subprocess, sys
rep = sys.argv[1]
# either an .ncu-rep or the CSV of its raw page (ncu -i REP --page raw --csv), which is what travels back from the GPU box
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if r]
start = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
rows = rows[start:]
hdr = rows[0]
rows = [rows[0]] + rows[2:]   # drop the units row
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        name = w.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '')
        print(f"{name[:52]:52s}", [r[i][:24] for r in rows[1:]])
