#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02t_pytest.log
timeout 300 python scripts/graph_timeline.py 2 3 plain > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02t_timeline_c2.txt; head -45 gpurun_out/r02t_timeline_c2.txt | cut -c1-160
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02t_bench_short.json 2> gpurun_out/r02t_bench_short.err; cat gpurun_out/r02t_bench_short.json
timeout 300 python bench.py --no-roofline --steps 64 --config 3 > gpurun_out/r02t_bench_short_c3.json 2> gpurun_out/r02t_bench_short_c3.err; cat gpurun_out/r02t_bench_short_c3.json
timeout 300 python scripts/perf_pointwise.py 64 > gpurun_out/r02t_perf_pointwise.log 2>&1; head -12 gpurun_out/r02t_perf_pointwise.log
