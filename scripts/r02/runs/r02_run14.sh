#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02o_pytest.log
timeout 300 python scripts/conv_census.py 2 > gpurun_out/r02o_conv_census.log 2>&1; head -30 gpurun_out/r02o_conv_census.log | cut -c1-200; tail -7 gpurun_out/r02o_conv_census.log
timeout 300 python scripts/graph_timeline.py 2 3 plain > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02o_timeline_c2.txt; head -30 gpurun_out/r02o_timeline_c2.txt | cut -c1-160
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02o_bench_short.json 2> gpurun_out/r02o_bench_short.err; cat gpurun_out/r02o_bench_short.json
