#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/perf_shapes.py > gpurun_out/r02p_perf_shapes.log 2>&1; cat gpurun_out/r02p_perf_shapes.log
