#!/bin/bash
set -x
mkdir -p gpurun_out
# 1. default bench line (what the driver runs) and the reference arm
timeout 900 python bench.py > gpurun_out/r02ab_bench_c2.json 2> gpurun_out/r02ab_bench_c2.err; echo "bench exit $?"; tail -c 600 gpurun_out/r02ab_bench_c2.err; head -c 1500 gpurun_out/r02ab_bench_c2.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r02ab_bench_reference.json 2> gpurun_out/r02ab_bench_reference.err; echo "ref exit $?"; cat gpurun_out/r02ab_bench_reference.json
# 2. launch list of the same command (eager so that every launch is a kernel launch; plain iterations only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02ab_launches.csv python bench.py --steps 2 --warmup 1 --eager --plain-only --no-roofline --no-cpu-baseline > gpurun_out/r02ab_ncu_launches.log 2>&1; tail -2 gpurun_out/r02ab_ncu_launches.log; wc -l gpurun_out/r02ab_launches.csv
# 3. full capture of the dominant tensor-core launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv -c 12 -o gpurun_out/r02ab_conv_full python scripts/ncu_targets.py 64 > gpurun_out/r02ab_ncu.log 2>&1; tail -3 gpurun_out/r02ab_ncu.log; ls -la gpurun_out/r02ab*.ncu-rep
