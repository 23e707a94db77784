#!/bin/bash
# Round-2 first GPU call for the kernels parked on branch wip/r02-unvalidated-kernels (none of them has run on
# a GPU yet): validate, then measure against the round-1 numbers in profiles/r01d_*.
#   gpurun --timeout 1200 -- 'bash scripts/r02_first_run.sh'
set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I textboxgan_b200/csrc -o gpurun_out/exp_halo \
     scripts/exp_halo_umma.cu textboxgan_b200/csrc/host_util.cu && timeout 120 gpurun_out/exp_halo 2>&1 | tee gpurun_out/r02_exp_halo.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest.log
# if the line above shows conv failures, these isolate the two staged epilogues (round-1 store paths):
TBG_IGEMM_STAGED=0 TBG_WGRAD_STAGED=0 TBG_IGEMM_MSUB=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r02_pytest_unstaged.log
TBG_LSTM_CLUSTER=1 timeout 300 python -m pytest tests -m gpu -q -k "lstm or train_step" 2>&1 | tail -5 | tee gpurun_out/r02_pytest_lstm_cluster.log
TBG_LSTM_CLUSTER=1 timeout 300 python scripts/step_timing.py 1 ocr graph noprof 2>&1 | tail -5
timeout 300 python scripts/perf_layers.py 32 2>&1 | tee gpurun_out/r02_layer_perf.log | tail -25
timeout 300 python scripts/step_timing.py 1 ocr graph noprof 2>&1 | tail -5
timeout 300 python scripts/graph_timeline.py 1 3 2>&1 | sed -n 3,30p
# last: a wrong addressing variant can trap the context
timeout 300 python scripts/perf_halo.py 32 2>&1 | tee gpurun_out/r02_perf_halo.log | tail -6
# if a halo variant works (v = 1: base_offset 0, v = 2: base_offset from the start address), the whole suite and the step
# with the 3x3 convolutions routed through it:
for v in 1 2; do TBG_CONV_HALO=$v timeout 600 python -m pytest tests -m gpu -q -k "train_step or fused or conv" 2>&1 | tail -3; done
