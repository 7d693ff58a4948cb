#!/usr/bin/env python
"""DRAM bytes and time per launch of every kernel in an ncu --set full report -> profiles/traffic.json
(read by bench.py for roofline.traffic).  Usage: ncu_traffic.py report.ncu-rep "source description" [out.json]"""
import csv, io, json, re, subprocess, sys
rep, source = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
num = lambda x: float(x.replace(",", ""))
kernels = {}
for r in data:
    name = re.sub(r"^(void )?(egs::)?", "", r[ki])
    name = re.split(r"[(]", name)[0].strip()
    e = kernels.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += num(r[ri]) * scale[units[ri]] + num(r[wi]) * scale[units[wi]]
    e["time_us"] += num(r[ti]) * tscale[units[ti]]
for e in kernels.values():
    e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
    e["time_us_per_launch"] = e["time_us"] / e["launches"]
json.dump({"source": source, "kernels": kernels}, open(out, "w"), indent=1)
for k, e in kernels.items():
    print(f"{k[:60]:60s} x{e['launches']:2d}  {e['dram_bytes_per_launch'] / 1e6:9.2f} MB  {e['time_us_per_launch']:9.1f} us")
