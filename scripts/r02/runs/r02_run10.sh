#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_small_kernels.py -m gpu -q -k "wgrad or second_order or halo" -s > gpurun_out/r02j_pytest.log 2>&1; tail -5 gpurun_out/r02j_pytest.log
timeout 300 python scripts/graph_timeline.py 3 3 plain 32 > gpurun_out/r02j_timeline_c3_b32.txt 2>&1; head -60 gpurun_out/r02j_timeline_c3_b32.txt
timeout 300 python scripts/sweep_halo.py 32 2>&1 | grep wgrad | tee gpurun_out/r02j_wgrad.log
