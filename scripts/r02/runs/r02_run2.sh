#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02b_pytest.log
timeout 300 python scripts/perf_conv.py 32 2>&1 | tee gpurun_out/r02b_perf_conv.log
timeout 300 python scripts/perf_layers.py 32 2>&1 | tee gpurun_out/r02b_layer_perf.log | tail -25
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c2.json 2> gpurun_out/r02b_bench_c2.err; tail -c 1500 gpurun_out/r02b_bench_c2.err; head -c 3000 gpurun_out/r02b_bench_c2.json
