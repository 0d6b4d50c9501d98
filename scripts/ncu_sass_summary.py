"""Summarise an ncu report's SASS page: instruction counts / stall samples per barrier-delimited phase and per opcode.
usage: ncu_sass_summary.py report.ncu-rep [warp_tiles]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and len(r) > 5: cur["rows"].append(r)
for k in ks[:1]:
    h = k["hdr"]; iI = h.index("Instructions Executed"); iS = h.index("# Samples")
    print(k["name"][:90])
    tot = sum(int(r[iI]) for r in k["rows"]); ts = sum(int(r[iS]) for r in k["rows"])
    print("total warp instructions %.1fM (%.1f per unit), samples %d" % (tot / 1e6, tot / unit, ts))
    seg = []; acc = [0, 0, 0]; start = 0; cat = collections.Counter()
    segcats = []
    for i, r in enumerate(k["rows"]):
        op = [o for o in r[1].split() if not o.startswith("@")][0]
        base = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) else op.split(".")[0]
        n = int(r[iI]); acc[0] += n; acc[1] += int(r[iS]); acc[2] += n if base in ("DFMA", "DMUL", "DADD") else 0
        cat[base] += n
        if op.startswith("BAR") or op.startswith("EXIT") or i == len(k["rows"]) - 1:
            seg.append((start, i, op, *acc)); segcats.append(cat); acc = [0, 0, 0]; start = i + 1; cat = collections.Counter()
    for s, c in zip(seg, segcats):
        if s[3] < tot * 0.002: continue
        print("sass %5d-%5d %-8s instr %7.1f/unit (%4.1f%%) fp64 %6.1f/unit  samples %4.1f%%" % (
            s[0], s[1], s[2][:8], s[3] / unit, 100 * s[3] / tot, s[5] / unit, 100 * s[4] / max(ts, 1)))
        print("      ", {o: round(n / unit, 1) for o, n in c.most_common(18)})
