#!/bin/bash
# 2 GPUs of one box: the driver's scaling launch
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02an_2gpu.json 2> gpurun_out/bench_r02an_2gpu.err
cut -c1-260 gpurun_out/bench_r02an_2gpu.json
timeout 600 python -m pytest tests/test_gpu_pagraph.py tests/test_kmer_counter.py -m gpu -x -q 2>&1 | tail -3
