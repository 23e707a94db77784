#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02ac_pytest.log | cut -c1-200
timeout 300 python scripts/perf_shapes.py > gpurun_out/r02ac_perf_shapes.log 2>&1; grep wgrad gpurun_out/r02ac_perf_shapes.log
timeout 300 python scripts/perf_layers.py 32 > gpurun_out/r02ac_layer_perf.log 2>&1; tail -22 gpurun_out/r02ac_layer_perf.log
timeout 300 python scripts/perf_layers.py 64 > gpurun_out/r02ac_layer_perf_b64.log 2>&1; tail -22 gpurun_out/r02ac_layer_perf_b64.log
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02ac_bench_short.json 2>/dev/null; cat gpurun_out/r02ac_bench_short.json
