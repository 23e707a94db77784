#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02h_pytest.log 2>&1; tail -8 gpurun_out/r02h_pytest.log
timeout 600 python bench.py --steps 32 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_c2.json 2> gpurun_out/r02h_bench_c2.err; tail -c 1500 gpurun_out/r02h_bench_c2.err; head -c 1200 gpurun_out/r02h_bench_c2.json
timeout 300 python scripts/graph_timeline.py 2 2 pl > gpurun_out/r02h_timeline_c2_pl.txt 2>&1; head -30 gpurun_out/r02h_timeline_c2_pl.txt
timeout 300 python scripts/graph_timeline.py 2 2 r1pl > gpurun_out/r02h_timeline_c2_r1pl.txt 2>&1; head -30 gpurun_out/r02h_timeline_c2_r1pl.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv -c 10 -o gpurun_out/r02h_conv_full python scripts/ncu_targets.py 64 > gpurun_out/r02h_ncu.log 2>&1; tail -3 gpurun_out/r02h_ncu.log; ls -la gpurun_out/*.ncu-rep
