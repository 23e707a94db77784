#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02ak_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02ak_smoke.log 2>&1; tail -3 gpurun_out/r02ak_smoke.log
timeout 900 python bench.py > gpurun_out/r02ak_bench_c2.json 2> gpurun_out/r02ak_bench_c2.err; echo "bench exit $?"; head -c 700 gpurun_out/r02ak_bench_c2.json
for v in plain pl r1pl; do timeout 300 python scripts/graph_timeline.py 2 2 $v > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02ak_timeline_c2_$v.txt; head -3 gpurun_out/r02ak_timeline_c2_$v.txt | cut -c1-200; done
