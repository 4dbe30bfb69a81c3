#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep: duration, DRAM bytes, DRAM %, occupancy, issue %: ncu_list.py <rep>"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr = rows[0]
keys = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "t"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "winst"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
for r in rows[2:]:
    print("  ".join(f"{n}={r[hdr.index(k)]}{rows[1][hdr.index(k)]}" for k, n in keys if k in hdr))
