#!/usr/bin/env python
"""Groups the SASS of one kernel of an ncu report into basic blocks by executed count and prints the instruction mix:
ncu_mix.py report.ncu-rep kernel_regex.  Blocks with the same executed count are one loop level, so this shows
where the warp instructions of a kernel go (per-batch code vs per-trip code vs prologue)."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, iex, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[isrc].strip(), int(r[iex]), int(r[ist] or 0)) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
total = sum(e for _, e, _ in data)
print(f"{len(data)} SASS instructions, {total/1e6:.1f} M warp instructions executed")
# group consecutive instructions with (nearly) equal executed counts
groups, cur = [], None
for i, (s, e, smp) in enumerate(data):
    if cur and cur["lo"] * 0.97 <= e <= cur["hi"] * 1.03 and e > 0:
        cur["n"] += 1; cur["sum"] += e; cur["smp"] += smp; cur["lo"] = min(cur["lo"], e); cur["hi"] = max(cur["hi"], e); cur["ops"].append(s.split()[0] if not s.startswith("@") else s.split()[1])
    else:
        cur = dict(start=i, n=1, sum=e, smp=smp, lo=e, hi=e, ops=[s.split()[0] if not s.startswith("@") else s.split()[1]])
        groups.append(cur)
for g in groups:
    if g["sum"] < total * 0.004:
        continue
    mix = {}
    for o in g["ops"]:
        k = o.split(".")[0]
        mix[k] = mix.get(k, 0) + 1
    top = " ".join(f"{k}:{v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:14])
    print(f"  @{g['start']:4d} n={g['n']:3d} exec/instr={g['sum']/g['n']/1e6:8.3f}M  share={100*g['sum']/total:5.1f}%  samples={g['smp']:6d} | {top}")
