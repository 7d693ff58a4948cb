#!/bin/bash
# one view per call: host enqueue vs device time, and the launch list of the pattern (metric and cfg2)
mkdir -p gpurun_out
for W in metric cfg2; do
  python scripts/c1_breakdown.py $W 2>&1 | tail -1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1_$W.csv python scripts/seq_views.py $W 2 2 > gpurun_out/launches_c1_$W.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_c1_$W.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
L=[(r[ki],float(r[vi].replace(',',''))) for r in rows[start+2:] if len(r)>vi and r[vi]]
idx=[i for i,(k,v) in enumerate(L) if 'projection_fwd' in k]
a,b=idx[-2],idx[-1]
tot=0
for k,v in L[a:b]:
    tot+=v/1000; print('  ',k[:60],round(v/1000,1))
print('  sum of kernels per view (us):',round(tot,1),'launches',b-a)
PY
done
