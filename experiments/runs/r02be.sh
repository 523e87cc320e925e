#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_index.py tests/test_host_binary.py -m gpu -x -q > gpurun_out/gpu_tests_r02be.log 2>&1
tail -3 gpurun_out/gpu_tests_r02be.log
timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02be.log 2>&1
tail -1 gpurun_out/seed_r02be.log | cut -c1-420
