#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_parity.py -m gpu -q > gpurun_out/r02y_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02y_pytest.log | cut -c1-300
timeout 300 python scripts/graph_timeline.py 3 3 plain 32 > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02y_timeline_c3_b32.txt; head -50 gpurun_out/r02y_timeline_c3_b32.txt | cut -c1-160
