#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02w_pytest.log
timeout 300 python scripts/graph_timeline.py 2 2 pl > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02w_timeline_c2_pl.txt; head -4 gpurun_out/r02w_timeline_c2_pl.txt | cut -c1-200
timeout 300 python scripts/graph_timeline.py 2 2 r1pl > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02w_timeline_c2_r1pl.txt; head -4 gpurun_out/r02w_timeline_c2_r1pl.txt | cut -c1-200
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02w_bench_short.json 2> gpurun_out/r02w_bench_short.err; cat gpurun_out/r02w_bench_short.json
timeout 300 python scripts/perf_pointwise.py 64 > gpurun_out/r02w_perf_pointwise.log 2>&1; grep fir4 gpurun_out/r02w_perf_pointwise.log
