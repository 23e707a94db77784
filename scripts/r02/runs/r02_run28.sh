#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -c 6 -o gpurun_out/r02aa_wgrad python scripts/ncu_wgrad.py > gpurun_out/r02aa_ncu.log 2>&1; tail -3 gpurun_out/r02aa_ncu.log; ls -la gpurun_out/r02aa*.ncu-rep
