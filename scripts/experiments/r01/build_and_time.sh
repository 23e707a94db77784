#!/bin/bash
# A/B: round-1 conv_igemm kernel (4 epilogue warps, direct stores, no sub-tiles) vs the current one, same process.
set -e
cd "$(dirname "$0")/../../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr \
  -I textboxgan_b200/csrc -o scripts/experiments/r01/libr01.so scripts/experiments/r01/conv_igemm_r01.cu textboxgan_b200/csrc/host_util.cu
# the current kernel with four epilogue warps (256 threads), to separate the thread count from the other changes
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared --expt-relaxed-constexpr -DTBG_IGEMM_EPI_WARPS=4 \
  -I textboxgan_b200/csrc -o scripts/experiments/r01/libepi4.so textboxgan_b200/csrc/conv_igemm.cu textboxgan_b200/csrc/conv_halo.cu textboxgan_b200/csrc/host_util.cu
python scripts/experiments/r01/time_ab.py
