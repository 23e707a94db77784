#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pairs" > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r02ae_pytest.log | cut -c1-300
timeout 300 python scripts/perf_cta2.py 64 > gpurun_out/r02ae_perf_cta2.log 2>&1; cat gpurun_out/r02ae_perf_cta2.log | cut -c1-300
