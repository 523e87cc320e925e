#!/bin/bash
# 4 GPUs of one box: the driver's scaling launch
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 --pagraph-reads 0 > gpurun_out/bench_r02bg_4gpu.json 2> gpurun_out/bench_r02bg_4gpu.err
cut -c1-260 gpurun_out/bench_r02bg_4gpu.json
