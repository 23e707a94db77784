#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 bash scripts/experiments/r01/build_and_time.sh 2>&1 | tee gpurun_out/r02f_igemm_ab.log
timeout 900 python -m pytest tests/test_gpu_step_parity.py tests/test_data_parallel_gloo.py -m gpu -q -s > gpurun_out/r02f_pytest_step.log 2>&1; tail -5 gpurun_out/r02f_pytest_step.log
timeout 300 python scripts/graph_timeline.py 2 2 pl > gpurun_out/r02f_timeline_c2_pl.txt 2>&1; head -40 gpurun_out/r02f_timeline_c2_pl.txt
timeout 300 python scripts/graph_timeline.py 2 2 r1pl > gpurun_out/r02f_timeline_c2_r1pl.txt 2>&1; head -40 gpurun_out/r02f_timeline_c2_r1pl.txt
