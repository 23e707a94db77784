#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02e_pytest.log 2>&1; tail -12 gpurun_out/r02e_pytest.log
timeout 300 python scripts/sweep_halo.py 32 2>&1 | tee gpurun_out/r02e_sweep_halo.log
timeout 300 bash scripts/experiments/r01/build_and_time.sh 2>&1 | tee gpurun_out/r02e_igemm_ab.log
timeout 300 python scripts/perf_pointwise.py 64 2>&1 | tee gpurun_out/r02e_perf_pointwise.log
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err; tail -c 1500 gpurun_out/r02e_bench_c2.err; head -c 1500 gpurun_out/r02e_bench_c2.json
timeout 300 python scripts/graph_timeline.py 2 3 > gpurun_out/r02e_timeline_c2.txt 2>&1; head -24 gpurun_out/r02e_timeline_c2.txt
