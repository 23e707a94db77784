#!/bin/bash
# launch list of the bench command at HEAD + per-kernel DRAM / tensor-pipe capture of one whole eager iteration
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02at_launches.csv python bench.py --steps 2 --warmup 1 --eager --plain-only --no-roofline --no-cpu-baseline > gpurun_out/r02at_ncu_launches.log 2>&1; tail -2 gpurun_out/r02at_ncu_launches.log; wc -l gpurun_out/r02at_launches.csv
M=$(python -c "import sys; sys.path.insert(0,'scripts'); import ncu_step_all as n; print(n.METRICS)")
timeout 1200 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02at_step_all.csv python scripts/ncu_step_all.py run 2 > gpurun_out/r02at_step_all.log 2>&1; tail -2 gpurun_out/r02at_step_all.log; wc -l gpurun_out/r02at_step_all.csv
