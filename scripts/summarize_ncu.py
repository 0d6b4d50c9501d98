#!/usr/bin/env python
"""Turns gpurun_out/launches_<tag>.csv (+ prof_<tag>.ncu-rep) into profiles/<tag>_launches.txt and <tag>_<kernel>.txt."""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else ""          # "mg_": the launch list of scripts/mg_ncu.sh
out = []
with open("gpurun_out/%slaunches_%s.csv" % (prefix, tag)) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
out.append("# command: see scripts/ncu_round.sh (bench.py --profiler-range under ncu --profile-from-start off: the launches of ONE timed step)"
           if not prefix else "# command: see scripts/mg_ncu.sh (python scripts/mg_probe.py 4000 nojac under ncu: two S16 solves through FEMSolver.Solve with the multigrid solver, set-up + 11 extra V-cycles)")
out.append("# ncu --metrics gpu__time_duration.sum --clock-control none  (tag %s); per-launch times are cold-cache and serialised: compare SHARES" % tag)
out.append("# launches  total_ms  avg_us  share  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%7d  %10.3f  %9.1f  %5.1f%%  %s" % (v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot, k))
open("profiles/%s_%slaunches.txt" % (tag, prefix), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
try:
    raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s.ncu-rep" % tag, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    o = ["# ncu --set full --clock-control none --import-source on  (tag %s); selected raw metrics per captured launch" % tag]
    for r in rows[2:]:
        o.append("kernel: " + r[hdr.index("Kernel Name")][:160])
        for w in WANT:
            if w in hdr:
                o.append("  %-80s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        if "dram__bytes_read.sum" in hdr:
            def val(name):
                x = float(r[hdr.index(name)].replace(",", "")); u = units[hdr.index(name)]
                return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            o.append("  traffic (dram read+write) = %.0f bytes" % (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")))
    open("profiles/%s_full.txt" % tag, "w").write("\n".join(o) + "\n")
    print("\n".join(o[:40]))
except Exception as e:
    print("no full report:", e)
