#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_step_parity.py 2>&1 | tail -30 | tee gpurun_out/r02c_pytest.log
timeout 1200 python -m pytest tests/test_gpu_step_parity.py -m gpu -q -s 2>&1 | tail -60 | tee gpurun_out/r02c_pytest_step.log
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err; tail -c 1500 gpurun_out/r02c_bench_c2.err; head -c 1200 gpurun_out/r02c_bench_c2.json
timeout 300 python scripts/graph_timeline.py 2 3 > gpurun_out/r02c_timeline_c2.txt 2>&1; head -45 gpurun_out/r02c_timeline_c2.txt
