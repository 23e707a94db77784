#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r02m_pytest.log
timeout 300 python scripts/conv_census.py 2 > gpurun_out/r02m_conv_census.log 2>&1; tail -90 gpurun_out/r02m_conv_census.log
