#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small_kernels.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02ai_pytest.log | cut -c1-200
timeout 300 python scripts/perf_pointwise.py 64 > gpurun_out/r02ai_perf_pointwise.log 2>&1; cat gpurun_out/r02ai_perf_pointwise.log
timeout 300 python scripts/perf_pointwise.py 32 > gpurun_out/r02ai_perf_pointwise_b32.log 2>&1; head -12 gpurun_out/r02ai_perf_pointwise_b32.log
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02ai_bench_short.json 2>/dev/null; cat gpurun_out/r02ai_bench_short.json | cut -c1-330
timeout 300 python bench.py --no-roofline --steps 64 --config 3 > gpurun_out/r02ai_bench_short_c3.json 2>/dev/null; cat gpurun_out/r02ai_bench_short_c3.json | cut -c1-330
