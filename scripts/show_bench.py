#!/usr/bin/env python
"""Prints value / e2e / per-stage times of bench.py JSON logs: show_bench.py log [log ...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "parse failed:", e); print(open(f).read()[-1500:]); continue
    ts = d.get("train_step", {}).get("it_per_s")
    print(f"{f}: value {d['value']:.1f} ms/step {d['ms_per_step']:.3f} e2e {d['e2e']['value']:.1f}" + (f" it/s {ts:.2f}" if ts else ""))
    if "stages" in d:
        print("   " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d["stages"].items()))
