#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02l_pytest.log
timeout 300 python scripts/sweep_halo.py > gpurun_out/r02l_sweep_halo.log 2>&1; tail -30 gpurun_out/r02l_sweep_halo.log
timeout 300 python scripts/perf_pointwise.py 64 > gpurun_out/r02l_perf_pointwise.log 2>&1; tail -30 gpurun_out/r02l_perf_pointwise.log
timeout 300 python scripts/graph_timeline.py 2 3 plain > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02l_timeline_c2.txt; head -40 gpurun_out/r02l_timeline_c2.txt
timeout 300 python scripts/graph_timeline.py 3 3 plain > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02l_timeline_c3_b32.txt; head -40 gpurun_out/r02l_timeline_c3_b32.txt
timeout 600 python bench.py > gpurun_out/r02l_bench_c2.json 2> gpurun_out/r02l_bench_c2.err; echo "bench exit $?"; head -c 3000 gpurun_out/r02l_bench_c2.json
