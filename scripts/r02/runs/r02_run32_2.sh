#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 64 --warmup 3 > gpurun_out/r02ad_bench_c3_n2.json 2> gpurun_out/r02ad_bench_c3_n2.err
echo "exit code $?"
grep -v NCCL gpurun_out/r02ad_bench_c3_n2.json | head -c 400
