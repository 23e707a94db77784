#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 16 --warmup 3 > gpurun_out/r02i_bench_c3_n2.json 2> gpurun_out/r02i_bench_c3_n2.err
echo "exit code $?"
tail -c 1200 gpurun_out/r02i_bench_c3_n2.err; head -c 1500 gpurun_out/r02i_bench_c3_n2.json
