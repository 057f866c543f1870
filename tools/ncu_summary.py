#!/usr/bin/env python3
"""Per-kernel summary of an ncu report (raw page): time, DRAM bytes, issue %, occupancy, registers.
usage: tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-70s %s %s" % (w, r[i], units[i]))
