"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv > X.csv):  python scripts/ncu_summary.py X.csv [pattern...]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
pats = sys.argv[2:] or [
    r"^gpu__time_duration.sum$", r"^dram__bytes_(read|write).sum$", r"gpu__dram_throughput.avg.pct", r"launch__registers_per_thread",
    r"^sm__warps_active.avg.pct", r"^smsp__inst_executed.sum$", r"^smsp__inst_executed.avg.per_cycle_active$", r"^sm__inst_executed_pipe_.*sum$",
    r"^smsp__inst_executed_pipe_.*sum$", r"pipe_fp64.*pct", r"bank_conflicts_pipe_lsu_mem_shared", r"wavefronts_mem_shared", r"^smsp__average_warp.*ratio$",
    r"^smsp__average_warps_issue_stalled.*per_warp_active.pct$", r"^sm__cycles_elapsed.max$", r"^smsp__cycles_active.avg$", r"shared_(ld|st).sum$",
    r"^smsp__issue_active.avg.pct", r"^sm__throughput.avg.pct", r"l1tex__throughput.avg.pct", r"^launch__", r"local_(load|store)",
]
for r in rows[2:]:
    print("=== kernel", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(re.search(p, h) for p in pats):
            v = r[i]
            if v not in ("", "0", "n/a"):
                print(f"  {h:95s} {units[i]:14s} {v}")
