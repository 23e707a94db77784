#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 6 -o gpurun_out/r02q_igemm_epi python scripts/ncu_igemm_epilogue.py > gpurun_out/r02q_ncu.log 2>&1; tail -3 gpurun_out/r02q_ncu.log; ls -la gpurun_out/*.ncu-rep
