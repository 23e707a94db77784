#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02g_pytest.log 2>&1; tail -8 gpurun_out/r02g_pytest.log
timeout 300 python scripts/perf_pointwise.py 64 2>&1 | tee gpurun_out/r02g_perf_pointwise.log | grep -E "64x64x256x128|64x32x128x128"
timeout 300 python scripts/perf_layers.py 32 2>&1 | tee gpurun_out/r02g_layer_perf.log | tail -22
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_c2.json 2> gpurun_out/r02g_bench_c2.err; tail -c 1500 gpurun_out/r02g_bench_c2.err; head -c 1500 gpurun_out/r02g_bench_c2.json
timeout 300 python scripts/graph_timeline.py 2 3 > gpurun_out/r02g_timeline_c2.txt 2>&1; head -24 gpurun_out/r02g_timeline_c2.txt
