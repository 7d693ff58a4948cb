#!/bin/bash
# Round 2, call M: full GPU suite on the tight-list route, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/r2m_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r2m_tests.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2m_bench.err
python scripts/show_bench.py gpurun_out/r2m_bench.log 2>/dev/null | cut -c1-330
