"""Top SASS lines of an ncu report by stall samples.  usage: ncu_hot.py report.ncu-rep [n]"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(txt.splitlines()))
hdr = None; body = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if body: break
        continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) > 5: body.append(r)
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
ts = sum(int(r[iS]) for r in body)
order = sorted(range(len(body)), key=lambda i: -int(body[i][iS]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[j]), hdr[j]) for j in stall), reverse=True)[:2]
    print("%5d %5.1f%% %-70s %s" % (i, 100 * int(r[iS]) / ts, r[1].strip()[:70], ", ".join("%s %d" % (b, a) for a, b in top if a)))
