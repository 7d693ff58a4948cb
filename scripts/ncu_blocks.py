#!/usr/bin/env python
"""Basic-block level instruction-count distribution of one kernel from an ncu report (source page).
Usage: ncu_blocks.py report.ncu-rep kernel-regex"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", sys.argv[3] if len(sys.argv) > 3 else "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw))]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source"); ex = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
tot_ex = sum(num(r[ex]) for r in data); tot_s = sum(num(r[si]) for r in data)
print("kernel:", rows[0][1][:90]); print("total warp-instructions", tot_ex, "stall samples", tot_s)
blocks = []; cur = None
for i, r in enumerate(data):
    e = num(r[ex])
    if cur is None or cur[0] != e:
        cur = [e, 0, i, 0, 0, r[src].strip()[:44]]; blocks.append(cur)
    cur[1] += 1; cur[3] += num(r[te]); cur[4] += num(r[si])
for e, n, i0, t, s, first in blocks:
    if e * n > 0.004 * tot_ex or s > 0.01 * tot_s:
        print(f"@{i0:4d} n={n:3d} exec={e:10d} inst_share={100*e*n/max(1,tot_ex):5.1f}% stall_share={100*s/max(1,tot_s):5.1f}% thr={t/max(1,e*n):5.1f}  {first}")
