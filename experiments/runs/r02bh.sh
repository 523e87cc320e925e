#!/bin/bash
# 8 GPUs of one box: the driver's scaling launch
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --pagraph-reads 0 > gpurun_out/bench_r02bh_8gpu.json 2> gpurun_out/bench_r02bh_8gpu.err
cut -c1-260 gpurun_out/bench_r02bh_8gpu.json
