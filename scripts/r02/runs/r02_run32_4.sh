#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 64 --warmup 3 > gpurun_out/r02ad_bench_c3_n4.json 2> gpurun_out/r02ad_bench_c3_n4.err
echo "exit code $?"
grep -v NCCL gpurun_out/r02ad_bench_c3_n4.json | head -c 400
