#!/usr/bin/env python3
"""Condense an .ncu-rep into the JSON summary kept under profiles/: usage  tools/ncu_summary.py REPORT.ncu-rep OUT.json "command line" [note]"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
note = sys.argv[4] if len(sys.argv) > 4 else ""
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__cluster_dim_x", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(txt.split("\n")))
h, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    if not r:
        continue
    m = {k: dict(unit=units[h.index(k)], value=r[h.index(k)]) for k in KEYS if k in h}
    res.append(dict(kernel=r[h.index("Kernel Name")], metrics=m))
json.dump(dict(command=cmd, note=note, launches=res), open(out, "w"), indent=1)
print(out, len(res), "launch(es)")
