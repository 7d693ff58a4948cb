#!/bin/bash
# Round 2, call Y: bound-based capacity check (no wait for the route), spare blocks leave before the ticket.
T=${1:-r2y}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
for W in metric cfg2; do python scripts/c1_breakdown.py $W 2>&1 | tail -1; done
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.log 2>/dev/null | grep -v "parity\|cpu_baseline\|roofline" | head -40 | cut -c1-300
