#!/usr/bin/env python
"""gpurun_out/prof_<tag>.ncu-rep -> profiles/<out>.txt (selected raw metrics per captured launch)."""
import csv
import subprocess
import sys

tag, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
W = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
     "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__grid_size",
     "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
     "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
o = ["# ncu --set full --clock-control none --import-source on (tag %s): selected raw metrics per captured launch" % tag]
for row in r[2:]:
    o.append("kernel: " + row[hdr.index("Kernel Name")][:140])
    for w in W:
        if w in hdr:
            o.append("  %-85s %s %s" % (w, row[hdr.index(w)], units[hdr.index(w)]))
    def val(name):
        x = float(row[hdr.index(name)].replace(",", "")); u = units[hdr.index(name)]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    o.append("  traffic (dram read+write) = %.0f bytes" % (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")))
open("profiles/%s.txt" % out, "w").write("\n".join(o) + "\n")
print("\n".join(o[:30]))
