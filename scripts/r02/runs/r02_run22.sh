#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python scripts/graph_timeline.py 2 2 pl > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02u_timeline_c2_pl.txt; head -60 gpurun_out/r02u_timeline_c2_pl.txt | cut -c1-160
timeout 300 python scripts/graph_timeline.py 2 2 r1pl > /dev/null 2>&1; cp gpurun_out/graph_timeline.txt gpurun_out/r02u_timeline_c2_r1pl.txt; head -60 gpurun_out/r02u_timeline_c2_r1pl.txt | cut -c1-160
