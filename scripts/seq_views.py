#!/usr/bin/env python
"""GPU box helper: the reference's one-view-per-call loop (bench.sequential_views) on one workload.
seq_views.py cfg2 [n_views] [steps] [fwd]"""
import json, sys
sys.path.insert(0, ".")
import torch
import bench
wl = sys.argv[1]
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 4
st = int(sys.argv[3]) if len(sys.argv) > 3 else 6
fwd = len(sys.argv) > 4 and sys.argv[4] == "fwd"
print(json.dumps(bench.sequential_views(wl, torch.device("cuda", 0), n_views=nv, steps=st, forward_only=fwd)))
