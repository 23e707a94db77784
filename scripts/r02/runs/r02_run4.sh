#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 bash scripts/experiments/r01/build_and_time.sh 2>&1 | tee gpurun_out/r02d_igemm_ab.log
timeout 1200 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_small_kernels.py -m gpu -q -s > gpurun_out/r02d_pytest_step.log 2>&1; tail -12 gpurun_out/r02d_pytest_step.log
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_c2.json 2> gpurun_out/r02d_bench_c2.err; tail -c 1500 gpurun_out/r02d_bench_c2.err; head -c 1500 gpurun_out/r02d_bench_c2.json
timeout 300 python scripts/graph_timeline.py 2 3 > gpurun_out/r02d_timeline_c2.txt 2>&1; head -24 gpurun_out/r02d_timeline_c2.txt
