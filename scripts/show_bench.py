#!/usr/bin/env python
"""Prints value / e2e / per-stage times of bench.py JSON logs: show_bench.py log [log ...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "parse failed:", e); print(open(f).read()[-1500:]); continue
    ts = d.get("train_step", {}).get("it_per_s")
    e2e = d.get("e2e", {}).get("value")
    print(f"{f}: value {d['value']:.1f} ms/step {d['ms_per_step']:.3f} e2e {e2e if e2e is None else round(e2e, 1)} launches {d.get('gpu_launches')}"
          + (f" it/s {ts:.2f}" if ts else ""))
    if "stages" in d:
        print("   " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d["stages"].items()))
        print("   standalone " + "  ".join(f"{k}={s['ms']:.3f}({s['frac']:.2f})" for k, s in d.get("stages_standalone", {}).items()))
    for k, v in d.get("call_pattern", {}).items():
        if isinstance(v, dict):
            print(f"   {k}: " + "  ".join(f"{a}={round(b, 3) if isinstance(b, float) else b}" for a, b in v.items() if a not in ("workload", "label", "mode")))
    for k in ("parity", "exchange_check", "cpu_baseline", "clocks", "roofline"):
        if k in d:
            print(f"   {k}: {json.dumps(d[k])[:600]}")
    for name, ph in (("step_phases", d.get("step_phases")), ("train_step.step_phases", d.get("train_step", {}).get("step_phases"))):
        if ph:
            r = lambda k: " ".join(f"{x[k]:.3f}" for x in ph["per_rank"])
            print(f"   {name}: render [{r('render_ms')}] exchange [{r('exchange_ms')}] optimizer [{r('optimizer_ms')}] ms; "
                  f"transfer seen by the slowest renderer (rank {ph['slowest_renderer']}): {ph['transfer_ms_seen_by_the_slowest_renderer']:.3f} ms")
