#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 32 --warmup 3 > gpurun_out/r02k_bench_c3_n8.json 2> gpurun_out/r02k_bench_c3_n8.err
echo "exit code $?"
tail -c 800 gpurun_out/r02k_bench_c3_n8.err; head -c 1200 gpurun_out/r02k_bench_c3_n8.json
