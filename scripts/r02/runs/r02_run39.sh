#!/bin/bash
# per-kernel DRAM / tensor-pipe capture of one whole eager iteration at HEAD (after the rgb_bwd ring and the FromRGB forward rewrite)
M=$(python -c "import sys; sys.path.insert(0,'scripts'); import ncu_step_all as n; print(n.METRICS)")
timeout 1200 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02az_step_all.csv python scripts/ncu_step_all.py run 2 > gpurun_out/r02az_step_all.log 2>&1; tail -2 gpurun_out/r02az_step_all.log; wc -l gpurun_out/r02az_step_all.csv
