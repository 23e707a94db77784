#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02aj_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r02aj_pytest.log | cut -c1-300
timeout 300 python bench.py --no-roofline --steps 64 > gpurun_out/r02aj_bench_short.json 2> gpurun_out/r02aj_bench_short.err; cat gpurun_out/r02aj_bench_short.json; tail -3 gpurun_out/r02aj_bench_short.err
timeout 300 python bench.py --no-roofline --steps 64 --config 3 > gpurun_out/r02aj_bench_short_c3.json 2> gpurun_out/r02aj_bench_short_c3.err; cat gpurun_out/r02aj_bench_short_c3.json
