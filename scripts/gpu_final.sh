#!/bin/bash
# End-of-round evidence on one GPU: full -m gpu suite, default bench line, smoke(), launch list of a step, ncu --set full
# of every library kernel of one late step, compute-sanitizer passes over scripts/sanitize_small.py.
T=${1:-r2final}
mkdir -p gpurun_out profiles/sanitizer
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider -rf > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench.err
python scripts/show_bench.py gpurun_out/${T}_bench.log 2>/dev/null | grep -v "parity\|cpu_baseline" | head -40 | cut -c1-330
BENCH="python bench.py --steps 2 --warmup 3 --quick --no-call-pattern --no-exchange-check"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv $BENCH > gpurun_out/launches_${T}.log 2>&1
echo "launch list rc=$?"
# 18 library kernels per step (projection fwd, visible, 2 x (histogram + passes), scan+emit, offsets, 2 schedule, blend fwd,
# blend bwd, projection bwd, stats); skip the first steps, keep one whole step and a bit
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"^(projection|visible|radix|isect|tile_|rasterize|densify)" -s 72 -c 20 -f -o gpurun_out/prof_${T} $BENCH > gpurun_out/prof_${T}.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_${T}.ncu-rep
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > profiles/sanitizer/${T}_$tool.txt 2>&1
  echo "$tool rc=$? $(grep -c '^ok' profiles/sanitizer/${T}_$tool.txt) ok-lines; $(grep 'ERROR SUMMARY' profiles/sanitizer/${T}_$tool.txt)"
done
cp profiles/sanitizer/${T}_*.txt gpurun_out/
