#!/usr/bin/env python
"""Print the binning chain (projection fwd .. blend fwd) of the last step in an ncu launch list (csv)."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
L = [(int(r[idi]), r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 2:] if len(r) > vi and r[vi]]
idx = [i for i, (a, k, v) in enumerate(L) if "projection_fwd" in k][-1]
tot = 0.0
for id_, k, v in L[idx:]:
    print(id_, k[:64], round(v / 1000, 1))
    if "rasterize_fwd" in k:
        break
    if "projection_fwd" not in k:
        tot += v / 1000
print("binning chain total (us):", round(tot, 1))
